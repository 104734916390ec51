"""Local solution pieces of a node: the polyhedral pieces of its solution map around the
current point.

Mirrors /root/reference/src/avi.jl:447-477 (process_solution_graph) and
/root/reference/src/avi_solutions.jl: comp_indices (:511-612, computed by
`qpn_comp_indices_batched` on the device), all_Ks (:200-215), local_piece (:400-496), expand
(:241-261), collect / expand_recipes! (:277-321), project_and_permute (:79-90).
"""
import itertools

import numpy as np

from . import polyhedra as ph
from .model import INF, Poly


def single_node_gavi(net, pid, polys, dec):
    """avi.jl:447-475: z = [x_dec; lambda], w = x_param,
    (Q_dd x_d + Q_dp w + q_d - A_d' lambda) comp. x_d free ; lambda comp. l <= A_d x_d + A_p w <= u."""
    n = net.n_vars
    par = [i for i in range(n) if i not in set(dec)]
    qp = net.qps[pid]
    if polys:
        AA = np.vstack([p.A for p in polys]); l2 = np.concatenate([p.l for p in polys]); u2 = np.concatenate([p.u for p in polys])
    else:
        AA, l2, u2 = np.zeros((0, n)), np.zeros(0), np.zeros(0)
    m, nd = len(l2), len(dec)
    g = dict(M=np.hstack([qp.Q[np.ix_(dec, dec)], -AA[:, dec].T]), N=qp.Q[np.ix_(dec, par)], o=qp.q[dec],
             l1=np.full(nd, -INF), u1=np.full(nd, INF), A=np.hstack([AA[:, dec], np.zeros((m, m))]), B=AA[:, par], l2=l2, u2=u2)
    return g, par


def all_Ks(mask):
    """avi_solutions.jl:200-215: every assignment of one admissible set (1..4, or 5..8 for the second
    block, encoded in the 4-bit masks) to every index.  A recipe is the tuple of chosen set ids."""
    choices = [[b + 1 for b in range(4) if (int(mk) >> b) & 1] for mk in mask]
    return set(itertools.product(*choices))


def local_piece(g, K):
    """avi_solutions.jl:400-496 with reducible_inds empty (as `expand` calls it): the polyhedron over
    (z, w) on which the complementarity pattern K holds.  K[i] in 1..4 (block-2 indices carry the same
    code; the reference's 5..8)."""
    d1, d2 = len(g["l1"]), len(g["l2"])
    n, m = d1 + d2, g["N"].shape[1]
    A = np.vstack([np.hstack([g["M"], g["N"]]),
                   np.hstack([np.zeros((d2, d1)), np.eye(d2), np.zeros((d2, m))]),
                   np.hstack([np.eye(d1), np.zeros((d1, d2)), np.zeros((d1, m))]),
                   np.hstack([g["A"], g["B"]])])
    lo, up = np.empty(2 * n), np.empty(2 * n)
    for i in range(n):
        k = K[i]
        if i < d1:
            o, l, u = g["o"][i], g["l1"][i], g["u1"][i]
            b = {1: (-o, INF, l, l), 2: (-o, -o, l, u), 3: (-INF, -o, u, u), 4: (-INF, INF, l, u)}[k]
        else:
            l, u = g["l2"][i - d1], g["u2"][i - d1]
            b = {1: (0.0, INF, l, l), 2: (0.0, 0.0, l, u), 3: (-INF, 0.0, u, u), 4: (-INF, INF, l, u)}[k]
        lo[i], up[i], lo[n + i], up[n + i] = b
    noisy = lo > up
    lo[noisy] = up[noisy]
    A = A.copy()
    A[np.abs(A) <= 1e-8] = 0.0                           # droptol!(A, 1e-8)
    meaningful = [r for r in range(2 * n) if (not np.isinf(lo[r]) or not np.isinf(up[r])) and np.any(A[r] != 0.0)]
    if not meaningful:
        return Poly(np.zeros((0, n + m)), [], [])
    return ph.simplify(Poly(A[meaningful], lo[meaningful], up[meaningful]))


def project_and_permute(piece, dec, par, n_vars, lp):
    """avi_solutions.jl:79-90: keep [z[0:nv]; w], then scatter the columns to x's ordering."""
    d, nv, npar = piece.dim, len(dec), len(par)
    keep = list(range(nv)) + list(range(d - npar, d))
    proj = ph.project(piece, keep, lp)
    A = np.zeros((len(proj), n_vars))
    A[:, dec] = proj.A[:, :nv]
    A[:, par] = proj.A[:, nv:]
    return ph.simplify(Poly(A, proj.l, proj.u))


def vertices_of_slice(piece, z, w, nv, lp, max_dim=6):
    """The get_verts call of `expand` (avi_solutions.jl:252-255): vertices of the multiplier polytope
    at the current primal point.  Only needed when exploration_vertices > 0; enumerated as the
    basic solutions of the sliced system (small dimensions only)."""
    n = len(z)
    dim = n - nv
    if dim == 0 or dim > max_dim:                 # decided by the sizes alone: no need to slice first
        return []
    fixed = {j: z[j] for j in range(nv)}
    fixed.update({n + j: w[j] for j in range(len(w))})
    S = ph.simplify(ph.poly_slice(piece, fixed))
    if len(S) == 0:
        return []
    rows = []
    for i in range(len(S)):
        if not np.isinf(S.l[i]):
            rows.append((S.A[i], S.l[i]))
        if not np.isinf(S.u[i]) and S.u[i] != S.l[i]:
            rows.append((S.A[i], S.u[i]))
    verts = []
    for comb in itertools.combinations(range(len(rows)), dim):
        Am = np.array([rows[k][0] for k in comb]); bm = np.array([rows[k][1] for k in comb])
        if abs(np.linalg.det(Am)) < 1e-9:
            continue
        v = np.linalg.solve(Am, bm)
        if ph.contains(S, v, tol=1e-6, closed=True) and not any(np.allclose(v, q, atol=1e-5) for q in verts):
            verts.append(v)
    return [np.concatenate([z[:nv], v, w]) for v in verts]


class LocalSolutions:
    """LocalGAVISolutions (avi_solutions.jl:92-129) + collect (:277-321)."""

    def __init__(self, engine, lp, g, z, w, dec, par, n_vars, max_vertices=0, cache=None, cache_key=None):
        self.engine, self.lp, self.g = engine, lp, g
        # pieces depend on (node, its constraint polys incl. the chosen child pieces, K) only -- not on the point:
        # memoised across the instances of a batch
        self.cache, self.cache_key = cache, cache_key
        self.z, self.w, self.dec, self.par, self.n_vars = np.asarray(z, float), np.asarray(w, float), list(dec), list(par), n_vars
        self.max_vertices = max_vertices
        self.unexplored_Ks = self._recipes(self.z, self.w)
        self.explored_Ks, self.polys = set(), []
        self._poly_keys = set()
        self.unexplored_vertices, self.explored_vertices = [], [self._vkey(np.concatenate([self.z, self.w]))]

    @staticmethod
    def _vkey(v):
        return tuple(np.round(v, 5) + 0.0)               # QuantizedVector, avi_solutions.jl:23-32

    def _recipes(self, z, w):
        mask = self.engine.comp_indices(self.g, z[None, :], w[None, :])[0]
        if (mask == 0).any():
            raise RuntimeError("comp_indices: an index belongs to no set (the reference's @assert)")
        return all_Ks(mask)

    def expand(self, K):
        """avi_solutions.jl:241-261."""
        cache = {} if self.cache is None else self.cache
        once = self.lp.once
        piece = once(cache, (self.cache_key, K, "piece"), lambda: local_piece(self.g, K))
        zw = np.concatenate([self.z, self.w])
        if len(piece) and ph.isempty(piece, self.lp, tol=1e-4, x=zw):
            return None, []
        verts = []
        if self.max_vertices > 0 and (len(piece) == 0 or ph.contains(piece, zw)):
            verts = vertices_of_slice(piece, self.z, self.w, len(self.dec), self.lp)
        proj = once(cache, (self.cache_key, K, "proj"), lambda: project_and_permute(piece, self.dec, self.par, self.n_vars, self.lp))
        return proj, verts

    def collect(self):
        while self.unexplored_Ks:
            for K in sorted(self.unexplored_Ks):          # deterministic order (the reference iterates a Set)
                piece, verts = self.expand(K)
                if piece is None:
                    continue
                if piece not in self._poly_keys:
                    self._poly_keys.add(piece); self.polys.append(piece)
                for v in verts:
                    key = self._vkey(v)
                    if key not in self.explored_vertices and all(key != self._vkey(q) for q in self.unexplored_vertices):
                        self.unexplored_vertices.append(v)
            self.explored_Ks |= self.unexplored_Ks
            self.unexplored_Ks = set()
            if not self.unexplored_vertices:
                break
            while self.unexplored_vertices and len(self.explored_vertices) < self.max_vertices:
                v = self.unexplored_vertices.pop()
                self.explored_vertices.append(self._vkey(v))
                n = len(self.z)
                self.unexplored_Ks |= self._recipes(v[:n], v[n:]) - self.explored_Ks
        return list(self.polys)


def process_solution_graph(net, pid, polys, dec, x, lam, engine, lp, exploration_vertices=0, cache=None):
    """avi.jl:447-477."""
    # keyed on the exact, ordered rows: lam rides in the row order of `polys`, so two lists that are equal as sets
    # (5-digit slice keys) but list their rows differently must not share a GAVI
    key = (pid, tuple(p.exact_key for p in polys))
    if cache is not None and ("gavi", key) in cache:
        g, par = cache[("gavi", key)]
    else:
        g, par = single_node_gavi(net, pid, polys, dec)
        if cache is not None:
            cache[("gavi", key)] = (g, par)
    z = np.concatenate([x[dec], lam])
    return LocalSolutions(engine, lp, g, z, x[par], dec, par, net.n_vars, max_vertices=exploration_vertices, cache=cache, cache_key=key)
