"""ORACLE (test infrastructure): dense numpy restatement of the reference's host
logic around the numeric hot path.  Every function cites the reference lines
it follows.  Only tests/, smoke() and bench.py's cpu_baseline may import it.

Not restated (third-party, absent): OSQP inside find_closest_feasible!
(/root/reference/src/avi.jl:79-99) -- the projection QP is solved exactly by
the same pivoting solve instead of ADMM to 1e-8; PATH inside solve_avi and in
the verify_solution fallback -- replaced by oracle.avi_pivot (see its header).
"""
import math
import numpy as np

from . import avi_pivot
from ._fma import fma
from .avi_pivot import SUCCESS, FAILURE

INF = math.inf


# --------------------------------------------------------------------------
# sets.jl: Slice / Poly
# --------------------------------------------------------------------------
def normalize_slice(a, l, u, rl=False, ru=False, tol=1e-8):
    """sets.jl:76-89.  rl/ru: True means strict '<'.  Returns (a,l,u,rl,ru)."""
    a = np.array(a, dtype=float)
    a[np.abs(a) <= tol] = 0.0                         # droptol!
    nrm = math.sqrt(float(np.dot(a, a)))
    if nrm <= tol:
        return np.zeros_like(a), l, u, rl, ru
    nz = np.nonzero(a)[0]
    lead = a[nz[0]]
    n = abs(lead)
    if lead >= 0:
        return a / n, l / n, u / n, rl, ru
    return -a / n, -u / n, -l / n, ru, rl


def _slice_key(a, l, u, rl, ru):
    """sets.jl:104-112: equality/hash on values rounded to 5 digits."""
    r5 = lambda v: (v if math.isinf(v) else round(v, 5)) + 0.0
    return (tuple(r5(x) for x in a), r5(l), r5(u), bool(rl), bool(ru))


class Poly:
    """BasicPoly (sets.jl:123-125,151-158) as stacked normalised slices.

    The reference iterates a Julia Set{Slice} (hash order); here
    rows keep first-insertion order after the same 5-digit de-duplication."""

    def __init__(self, A, l, u, rl=None, ru=None, normalize=True):
        A = np.atleast_2d(np.asarray(A, dtype=float))
        m = len(l)
        rl = [False] * m if rl is None else list(rl)
        ru = [False] * m if ru is None else list(ru)
        rows, seen = [], set()
        for i in range(m):
            s = normalize_slice(A[i], float(l[i]), float(u[i]), rl[i], ru[i]) if normalize \
                else (A[i].copy(), float(l[i]), float(u[i]), rl[i], ru[i])
            k = _slice_key(*s)
            if k in seen:
                continue
            seen.add(k)
            rows.append(s)
        d = A.shape[1]
        self.A = np.array([r[0] for r in rows]).reshape(len(rows), d)
        self.l = np.array([r[1] for r in rows], dtype=float)
        self.u = np.array([r[2] for r in rows], dtype=float)
        self.rl = np.array([r[3] for r in rows], dtype=bool)
        self.ru = np.array([r[4] for r in rows], dtype=bool)

    def __len__(self):
        return len(self.l)

    @property
    def dim(self):
        return self.A.shape[1]

    def keyset(self):
        return frozenset(_slice_key(self.A[i], self.l[i], self.u[i], self.rl[i], self.ru[i])
                         for i in range(len(self)))


def in_slice(x, a, l, u, rl, ru, tol):
    """sets.jl:850-853:  rl(l - tol, a'x) && ru(a'x - tol, u)."""
    ax = 0.0
    for j in range(len(a)):
        ax = fma(a[j], x[j], ax)
    lo_ok = (l - tol < ax) if rl else (l - tol <= ax)
    up_ok = (ax - tol < u) if ru else (ax - tol <= u)
    return lo_ok and up_ok


def in_poly(x, P, tol=1e-6):
    """sets.jl:820-825 (the length(x)==embedded_dim branch)."""
    assert len(x) == P.dim
    return all(in_slice(x, P.A[i], P.l[i], P.u[i], P.rl[i], P.ru[i], tol) for i in range(len(P)))


def net_polys(net):
    """add_constraint! stores Poly(A, lb-vals, ub-vals) (programs.jl:164)."""
    if not hasattr(net, "_polys"):
        net._polys = {cid: Poly(c["A"], c["l"], c["u"]) for cid, c in net.cons.items()}
    return net._polys


# --------------------------------------------------------------------------
# avi.jl: GAVI assembly
# --------------------------------------------------------------------------
def create_labeled_gavi_from_qp(net, pid, S):
    """avi.jl:205-251.  S: child id -> Poly.  Returns M1, q1, M2, l2, u2, dvars."""
    dvars = net.decision_inds(pid)
    n = len(dvars)
    qp = net.qps[pid]
    n_total = net.n_vars
    polys = net_polys(net)
    blocks = [polys[ci] for ci in qp["cons"]]
    A_i = np.vstack([p.A for p in blocks]) if blocks else np.zeros((0, n_total))
    l_i = np.concatenate([p.l for p in blocks]) if blocks else np.zeros(0)
    u_i = np.concatenate([p.u for p in blocks]) if blocks else np.zeros(0)
    kids = [S[j] for j in net.edges[pid]]
    A_S = np.vstack([p.A for p in kids]) if kids else np.zeros((0, n_total))
    l_S = np.concatenate([p.l for p in kids]) if kids else np.zeros(0)
    u_S = np.concatenate([p.u for p in kids]) if kids else np.zeros(0)
    M1 = np.hstack([qp["Q"][dvars, :], np.zeros((n, n)), -A_i[:, dvars].T, -A_S[:, dvars].T])
    q1 = qp["q"][dvars]
    M2 = np.vstack([A_i, A_S])
    return dict(dvars=dvars, M1=M1, q1=q1, M2=M2, l2=np.concatenate([l_i, l_S]),
                u2=np.concatenate([u_i, u_S]))


def combine_gavis(n, dec_inds, param_inds, lg):
    """avi.jl:305-377.  z = [dec; xi per player; lambda/psi per player]."""
    nd = len(dec_inds)
    pool = sorted(lg)
    xi_dim = {p: lg[p]["M1"].shape[0] for p in pool}
    lam_dim = {p: lg[p]["M1"].shape[1] - n - xi_dim[p] for p in pool}
    total_xi = sum(xi_dim.values())
    total_dual = total_xi + sum(lam_dim.values())
    xi_off, lam_off = {}, {}
    o1, o2 = 0, total_xi
    for p in pool:
        xi_off[p], lam_off[p] = o1, o2
        o1 += xi_dim[p]
        o2 += lam_dim[p]
    Ms, Ns, qs = [], [], []
    for p in pool:
        M1 = lg[p]["M1"]
        Mi = np.zeros((M1.shape[0], nd + total_dual))
        Mi[:, :nd] = M1[:, dec_inds]
        Mi[:, nd + xi_off[p]: nd + xi_off[p] + xi_dim[p]] = M1[:, n:n + xi_dim[p]]
        Mi[:, nd + lam_off[p]: nd + lam_off[p] + lam_dim[p]] = M1[:, n + xi_dim[p]:]
        Ms.append(Mi)
        Ns.append(M1[:, param_inds])
        qs.append(lg[p]["q1"])
    A = np.vstack([lg[p]["M2"][:, dec_inds] for p in pool])
    B = np.vstack([lg[p]["M2"][:, param_inds] for p in pool])
    l2 = np.concatenate([lg[p]["l2"] for p in pool])
    u2 = np.concatenate([lg[p]["u2"] for p in pool])
    top_M = np.zeros((nd, nd + total_dual))
    for p in pool:
        for di, d in enumerate(dec_inds):
            if d in lg[p]["dvars"]:
                top_M[di, nd + xi_off[p] + lg[p]["dvars"].index(d)] = 1.0
    M = np.vstack([top_M] + Ms)
    N = np.vstack([np.zeros((nd, len(param_inds)))] + Ns)
    o = np.concatenate([np.zeros(nd)] + qs)
    A = np.hstack([A, np.zeros((A.shape[0], total_dual))])
    d1 = len(o)
    return dict(M=M, N=N, o=o, l1=np.full(d1, -INF), u1=np.full(d1, INF), A=A, B=B, l2=l2, u2=u2)


def convert(g):
    """avi.jl:113-128: GAVI -> AVI of size d1 + 2 d2."""
    d1, d2 = len(g["l1"]), len(g["l2"])
    npar = g["N"].shape[1]
    M = np.block([[g["M"], np.zeros((d1, d2))],
                  [g["A"], -np.eye(d2)],
                  [np.zeros((d2, d1)), np.eye(d2), np.zeros((d2, d2))]])
    N = np.vstack([g["N"], g["B"], np.zeros((d2, npar))])
    o = np.concatenate([g["o"], np.zeros(2 * d2)])
    l = np.concatenate([g["l1"], np.full(d2, -INF), g["l2"]])
    u = np.concatenate([g["u1"], np.full(d2, INF), g["u2"]])
    return dict(M=M, N=N, o=o, l=l, u=u)


def matvec(Mx, v):
    """Row-wise sequential fma accumulation: the summation order the C oracle
    and the CUDA kernels use, so results agree bit for bit."""
    out = np.zeros(Mx.shape[0])
    for i in range(Mx.shape[0]):
        acc = 0.0
        for j in range(Mx.shape[1]):
            acc = fma(Mx[i, j], v[j], acc)
        out[i] = acc
    return out


def find_closest_feasible(g, z0, w, solver=avi_pivot.solve_avi):
    """avi.jl:79-99:  min 0.5|z - z0|^2  s.t.  l2 - Bw <= A z <= u2 - Bw.

    Only the columns of A that are not structurally zero can move, so the KKT
    system  [I -A'; A 0] (+ slack rows as in convert)  is formed over those.
    Returns the projected z0 (unchanged when already feasible)."""
    A, l2, u2 = g["A"], g["l2"], g["u2"]
    d2 = len(l2)
    if d2 == 0:
        return z0.copy(), 0
    c = matvec(g["B"], w)
    s0 = matvec(A, z0) + c
    if all(l2[i] <= s0[i] <= u2[i] for i in range(d2)):
        return z0.copy(), 0
    cols = [j for j in range(A.shape[1]) if np.any(A[:, j] != 0.0)]
    k = len(cols)
    Ac = A[:, cols]
    n = k + 2 * d2
    M = np.zeros((n, n))
    M[:k, :k] = np.eye(k)
    M[:k, k:k + d2] = -Ac.T
    M[k:k + d2, :k] = Ac
    M[k:k + d2, k + d2:] = -np.eye(d2)
    M[k + d2:, k:k + d2] = np.eye(d2)
    rest = matvec(A, z0) - matvec(Ac, z0[cols])
    q = np.concatenate([-z0[cols], rest + c, np.zeros(d2)])
    l = np.concatenate([np.full(k + d2, -INF), l2])
    u = np.concatenate([np.full(k + d2, INF), u2])
    start = np.concatenate([z0[cols], np.zeros(d2), s0])
    z, status, piv, _ = solver(M, q, l, u, start)
    out = z0.copy()
    if status == SUCCESS:
        out[cols] = z[:k]
    return out, piv


def solve_gavi(g, z0, w, presolve=True, solver=avi_pivot.solve_avi):
    """avi.jl:101-111."""
    piv0 = 0
    if presolve:
        z0, piv0 = find_closest_feasible(g, z0, w, solver)
    avi = convert(g)
    d1, d2 = len(g["l1"]), len(g["l2"])
    s = matvec(g["A"], z0) + matvec(g["B"], w)
    z0s = np.concatenate([z0, s])
    q = matvec(avi["N"], w) + avi["o"]
    z, status, piv, basis = solver(avi["M"], q, avi["l"], avi["u"], z0s)
    return dict(z=z[:d1 + d2], status=status, pivots=piv + piv0, basis=basis, z_full=z)


def level_gavi(net, players, S):
    """avi.jl:394-400."""
    dec = sorted(set().union(*[set(net.decision_inds(p)) for p in players]))
    par = [i for i in range(net.n_vars) if i not in dec]
    lg = {p: create_labeled_gavi_from_qp(net, p, S) for p in players}
    return combine_gavis(net.n_vars, dec, par, lg), dec, par


def solve_qep(net, players, x, S, solver=avi_pivot.solve_avi):
    """avi.jl:382-444.  Returns (x_opt or None, info)."""
    g, dec, par = level_gavi(net, players, S)
    w = x[par]
    z0 = np.concatenate([x[dec], np.zeros(g["M"].shape[1] - len(dec))])
    ret = solve_gavi(g, z0, w, solver=solver)
    if ret["status"] != SUCCESS:
        return None, ret
    x_opt = x.copy()
    x_opt[dec] = ret["z"][:len(dec)]
    return x_opt, ret


# --------------------------------------------------------------------------
# avi_solutions.jl: comp_indices
# --------------------------------------------------------------------------
def _approx(a, b, atol):
    """Julia isapprox(a,b;atol) on scalars with rtol=0 (infinite equal values match)."""
    if a == b:
        return True
    return abs(a - b) <= atol


def comp_indices_block(l, u, r, z, tol=1e-2):
    """avi_solutions.jl:511-562 with no requests: 4-bit mask per index,
    bit0 -> 1, bit1 -> 2, bit2 -> 3, bit3 -> 4."""
    out = np.zeros(len(z), dtype=np.int8)
    for i in range(len(z)):
        eq = _approx(l[i], u[i], tol)
        m = 0
        if _approx(z[i], l[i], tol) and r[i] >= -tol and not eq:
            m |= 1
        if (l[i] - tol <= z[i] <= u[i] + tol) and _approx(r[i], 0.0, tol) and not eq:
            m |= 2
        if _approx(z[i], u[i], tol) and r[i] <= tol and not eq:
            m |= 4
        if m == 0:
            assert eq, "comp_indices: index is in no set"
            m = 8
        out[i] = m
    return out


def comp_indices(g, z, w, tol=1e-2):
    """avi_solutions.jl:587-612: masks for the d1 block (values 1..4) then the
    d2 block (the reference's 5..8, stored as the same 4 bits)."""
    d1 = len(g["o"])
    r1 = matvec(g["M"], z) + matvec(g["N"], w) + g["o"]
    J1 = comp_indices_block(g["l1"], g["u1"], r1, z[:d1], tol)
    r2 = z[d1:]
    s2 = matvec(g["A"], z) + matvec(g["B"], w)
    J2 = comp_indices_block(g["l2"], g["u2"], r2, s2, tol)
    return np.concatenate([J1, J2])


# --------------------------------------------------------------------------
# qp_processing.jl: verify_solution
# --------------------------------------------------------------------------
def lstsq_basic(Abar, rhs, rank_tol=1e-10):
    """Least-squares  Abar * lam ~ rhs  by Householder QR with column pivoting;
    dependent columns get a zero multiplier (a basic solution, as SuiteSparseQR
    returns for `\\` at qp_processing.jl:115)."""
    A = np.array(Abar, dtype=float)
    b = np.array(rhs, dtype=float)
    m, k = A.shape
    perm = list(range(k))
    rank = 0
    for c in range(min(m, k)):
        norms = [math.sqrt(sum(A[i, j] * A[i, j] for i in range(c, m))) for j in range(c, k)]
        jmax = max(range(len(norms)), key=lambda t: (norms[t], -t))
        if norms[jmax] <= rank_tol:
            break
        j = c + jmax
        if j != c:
            A[:, [c, j]] = A[:, [j, c]]
            perm[c], perm[j] = perm[j], perm[c]
        alpha = norms[jmax]
        if A[c, c] > 0:
            alpha = -alpha
        v = A[c:, c].copy()
        v[0] -= alpha
        vn = float(np.dot(v, v))
        if vn > 0:
            for j2 in range(c, k):
                s = 2.0 * float(np.dot(v, A[c:, j2])) / vn
                A[c:, j2] -= s * v
            s = 2.0 * float(np.dot(v, b[c:])) / vn
            b[c:] -= s * v
        rank += 1
    lam_p = np.zeros(k)
    for i in range(rank - 1, -1, -1):
        acc = b[i]
        for j in range(i + 1, rank):
            acc -= A[i, j] * lam_p[j]
        lam_p[i] = acc / A[i, i]
    lam = np.zeros(k)
    for i in range(k):
        lam[perm[i]] = lam_p[i]
    return lam


def verify_solution(Qd, qd, A, l, u, dec, x, polys=None, tol=1e-4, solver=avi_pivot.solve_avi):
    """qp_processing.jl:57-149.  Qd = Q[dec,:], qd = q[dec]; A,l,u stacked rows of
    all constraint polys (feasibility at tol 1e-3 is tested per poly, :86).
    Returns (solution, lam or None, info)."""
    qt = matvec(Qd, x) + qd
    m = A.shape[0]
    ax = matvec(A, x) if m else np.zeros(0)
    if polys is None:
        feasible = all((l[i] - 1e-3 <= ax[i]) and (ax[i] - 1e-3 <= u[i]) for i in range(m))
    else:
        feasible = all(in_poly(x, P, tol=1e-3) for P in polys)
    if not feasible:
        return False, None, "infeasible"
    nrm = lambda v: math.sqrt(sum(t * t for t in v))
    if m == 0:
        return (nrm(qt) <= tol), (np.zeros(0) if nrm(qt) <= tol else None), "unconstrained"
    pos = [ax[i] < l[i] + 1e-2 for i in range(m)]
    neg = [ax[i] > u[i] - 1e-2 for i in range(m)]
    both = [pos[i] and neg[i] for i in range(m)]
    pos = [pos[i] and not both[i] for i in range(m)]
    neg = [neg[i] and not both[i] for i in range(m)]
    ip = [i for i in range(m) if pos[i]]
    ineg = [i for i in range(m) if neg[i]]
    ib = [i for i in range(m) if both[i]]
    Ad = A[:, dec]
    Abar = np.hstack([Ad[ip].T, -Ad[ineg].T, Ad[ib].T]) if (ip or ineg or ib) else np.zeros((len(dec), 0))
    lam = lstsq_basic(Abar, qt)
    res = matvec(Abar, lam) - qt if Abar.shape[1] else -qt
    ok = all(lam[t] > -tol for t in range(len(ip) + len(ineg))) and nrm(res) <= tol
    if ok:
        out = np.zeros(m)
        for t, i in enumerate(ip):
            out[i] = lam[t]
        for t, i in enumerate(ineg):
            out[i] = -lam[len(ip) + t]
        for t, i in enumerate(ib):
            out[i] = lam[len(ip) + len(ineg) + t]
        return True, out, "lsq"
    # fallback (:129-146): sign-constrained least squares through the AVI solve
    lb = np.array([-INF if (neg[i] or both[i]) else 0.0 for i in range(m)])
    ub = np.array([INF if (pos[i] or both[i]) else 0.0 for i in range(m)])
    G = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            acc = 0.0
            for t in range(len(dec)):
                acc = fma(Ad[i, t], Ad[j, t], acc)
            G[i, j] = acc
    h = -matvec(Ad, qt)
    lam2, status, _, _ = solver(G, h, lb, ub, np.zeros(m))
    if status != SUCCESS:
        return False, None, "dual solve failed"
    res = matvec(Ad.T.copy(), lam2) - qt
    if nrm(res) <= 1e-4:
        return True, lam2, "nnls"
    return False, lam2, "suboptimal"


# --------------------------------------------------------------------------
# algorithm.jl: one level without children (bottom level / flat Nash game)
# --------------------------------------------------------------------------
def node_view(net, pid, S=None):
    """What verify_solution reads for a node (qp_processing.jl:57-66): Q[dec,:], q[dec],
    stacked rows of its constraint polys followed by the chosen child pieces."""
    S = S or {}
    dec = net.decision_inds(pid)
    qp = net.qps[pid]
    polys = [net_polys(net)[ci] for ci in qp["cons"]] + [S[j] for j in net.edges[pid]]
    A = np.vstack([p.A for p in polys]) if polys else np.zeros((0, net.n_vars))
    l = np.concatenate([p.l for p in polys]) if polys else np.zeros(0)
    u = np.concatenate([p.u for p in polys]) if polys else np.zeros(0)
    return qp["Q"][dec, :], qp["q"][dec], A, l, u, np.array(dec, dtype=np.int32)


def projections_equal(a, b):
    """`proj_vals ≈ prev` (algorithm.jl:24) = isapprox with rtol = sqrt(eps), atol = 0."""
    nrm = lambda v: math.sqrt(sum(fma(t, t, 0.0) if False else t * t for t in v))
    dd = na = nb = 0.0
    for x, y in zip(a, b):
        e = x - y
        dd = fma(e, e, dd); na = fma(x, x, na); nb = fma(y, y, nb)
    return math.sqrt(dd) <= 1.4901161193847656e-8 * max(math.sqrt(na), math.sqrt(nb))


def solve_level_bottom(net, level, x_init, proj=None, max_iters=None):
    """algorithm.jl:13-118 for a level whose players have no children, with the C oracle
    doing the numerics.  proj: (nproj, nv) projection vectors or None."""
    from . import cport
    players = net.depth[level]
    views = [node_view(net, p) for p in players]
    g, dec, par = level_gavi(net, players, {})
    max_iters = max_iters or net.options["max_iters"]
    x = np.array(x_init, dtype=float)
    hist, piv, it = [], 0, 0
    for it in range(1, max_iters + 1):
        if proj is not None and len(proj):
            pv = [float(sum_fma(x, v)) for v in proj]
            if any(projections_equal(pv, h) for h in hist):
                return dict(solved=False, x=x, iters=it, pivots=piv, why="cycle")
            hist.append(pv)
        all_sol, lams = True, []
        for (Qd, qd, A, l, u, d) in views:
            sol, lam, how, act = cport.verify_solution(Qd, qd, A, l, u, d, x)
            piv += cport.verify_solution.last_fallback_pivots      # the fused kernel counts these too
            lams.append(lam if sol else np.zeros(len(l)))
            all_sol &= sol
        if all_sol:
            return dict(solved=True, x=x, iters=it, pivots=piv, lam=np.concatenate(lams) if lams else np.zeros(0))
        w = x[par]
        z0 = np.concatenate([x[dec], np.zeros(g["M"].shape[1] - len(dec))])
        ret = cport.gavi_solve(g, z0, w)
        piv += ret["pivots"]
        if ret["status"] != SUCCESS:
            return dict(solved=False, x=x, iters=it, pivots=piv, why="avi")
        xn = x.copy()
        xn[dec] = ret["z"][:len(dec)]
        dn = 0.0
        for a, b in zip(xn, x):
            e = a - b
            dn = fma(e, e, dn)
        if math.sqrt(dn) < 1e-4:
            return dict(solved=False, x=x, iters=it, pivots=piv, why="disagreement")
        x = xn
    return dict(solved=False, x=x, iters=it, pivots=piv, why="max_iters")


def sum_fma(x, v):
    acc = 0.0
    for a, b in zip(x, v):
        acc = fma(a, b, acc)
    return acc
