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


def multiplier_vertices(g, z, w, max_new, max_active=12, max_nd=8):
    """The get_verts call of `expand` (avi_solutions.jl:252-255, sets.jl:439-453): vertices of the multiplier polytope of
    a node at the current primal point.  Slicing a local piece at the primal part of z and at w leaves a polyhedron in
    the multipliers alone -- stationarity A_d' lam = qt, a sign per multiplier whose row is active, zero elsewhere; the
    recipe with the fewest "inactive" choices slices to the whole polytope Lambda(x), every other one to a face of it,
    so the vertices `collect` can queue are those of Lambda(x).  The rows fixed at zero are eliminated and the basic
    solutions of the remaining system enumerated (bases in lexicographic order); the reference uses Polyhedra.jl's
    double description.  Returns at most max_new vertices that differ from lam itself at 5 digits, as full z vectors."""
    d1, m = len(g["l1"]), len(g["l2"])
    z, w = np.asarray(z, float), np.asarray(w, float)
    lam = z[d1:]
    if max_new <= 0 or d1 > max_nd:                            # the caps of the native enumeration (csrc/net/vertex_enum.h)
        return []
    max_new = min(max_new, 15)
    Ad = g["A"][:, :d1]                                       # m x nd
    ax = Ad @ z[:d1] + g["B"] @ w
    qt = g["M"][:, :d1] @ z[:d1] + g["N"] @ w + g["o"]         # Q_dd x_d + Q_dp w + q_d
    lo = np.abs(ax - g["l2"]) <= 1e-6
    up = np.abs(ax - g["u2"]) <= 1e-6
    act = np.flatnonzero(lo | up)
    if (np.abs(np.delete(lam, act)) > 1e-6).any() or len(act) == 0 or len(act) > max_active:
        return []                                             # the point is in no piece at the slice tolerance / nothing to enumerate
    sgn = np.where(lo[act] & up[act], 0, np.where(lo[act], 1, -1))
    G = Ad[act].T                                             # nd x a
    if (np.abs(G @ lam[act] - qt) > 1e-6).any():
        return []
    # independent equations: elimination with full pivoting (as the native code picks them)
    W = G.copy(); rows, cols = list(range(d1)), list(range(len(act))); r = 0
    while r < min(d1, len(act)):
        sub = np.abs(W[np.ix_(rows[r:], cols[r:])])
        if sub.size == 0 or sub.max() <= 1e-9:
            break
        e, k = np.unravel_index(int(np.argmax(sub)), sub.shape)      # first maximum in row-major order
        rows[r], rows[r + e] = rows[r + e], rows[r]; cols[r], cols[r + k] = cols[r + k], cols[r]
        for ee in rows[r + 1:]:
            f = W[ee, cols[r]] / W[rows[r], cols[r]]
            if f != 0.0:
                W[ee, cols[r:]] -= f * W[rows[r], cols[r:]]
        r += 1
    a = len(act)
    if a <= r or int((sgn == 0).sum()) > r:
        return []
    eqs = sorted(rows[:r])
    free = set(np.flatnonzero(sgn == 0).tolist())
    key = lambda v: tuple(np.rint(np.asarray(v) * 1e5))
    seen, out = {key(lam[act])}, []
    for comb in itertools.combinations(range(a), r):
        if not free <= set(comb):
            continue
        Mx = G[np.ix_(eqs, comb)]
        if abs(np.linalg.det(Mx)) < 1e-12:
            continue
        try:
            y = np.linalg.solve(Mx, qt[eqs])
        except np.linalg.LinAlgError:
            continue
        cand = np.zeros(a); cand[list(comb)] = y
        if ((sgn != 0) & (sgn * cand < -1e-6)).any() or (np.abs(G @ cand - qt) > 1e-6).any():
            continue
        if key(cand) in seen:
            continue
        seen.add(key(cand))
        lv = np.zeros(m); lv[act] = cand
        out.append(np.concatenate([z[:d1], lv]))
        if len(out) == max_new:
            break
    return out


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
        """avi_solutions.jl:241-261 (the vertices of the slice are those of the node's multiplier polytope whatever K is:
        `collect` computes them once)."""
        cache = {} if self.cache is None else self.cache
        once = self.lp.once
        piece = once(cache, (self.cache_key, K, "piece"), lambda: local_piece(self.g, K))
        zw = np.concatenate([self.z, self.w])
        if len(piece) and ph.isempty(piece, self.lp, tol=1e-4, x=zw):
            return None
        return once(cache, (self.cache_key, K, "proj"), lambda: project_and_permute(piece, self.dec, self.par, self.n_vars, self.lp))

    def collect(self):
        """avi_solutions.jl:277-321: the recipes of the point, then -- exploration_vertices permitting -- those of the
        vertices of its multiplier polytope that are new, each group in sorted order (the reference iterates Sets)."""
        rounds = [sorted(self.unexplored_Ks)]
        explored = set(rounds[0])
        if self.max_vertices > 1:
            fresh = set()
            for v in multiplier_vertices(self.g, self.z, self.w, self.max_vertices - 1):
                fresh |= self._recipes(v, self.w) - explored
            rounds.append(sorted(fresh))
        for Ks in rounds:
            for K in Ks:
                piece = self.expand(K)
                if piece is None:
                    continue
                if piece not in self._poly_keys:
                    self._poly_keys.add(piece); self.polys.append(piece)
        self.explored_Ks, self.unexplored_Ks = explored | set(rounds[-1]), set()
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
