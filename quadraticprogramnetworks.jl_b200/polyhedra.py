"""Polyhedral set operations of the hot path's callers:
membership, emptiness / exemplar, subset tests, complement, intersection, projection,
simplification -- mirroring /root/reference/src/sets.jl.

Every LP / QP the reference hands to OSQP (sets.jl:387-397, 611-619; avi.jl:80-93) is solved here
by the GPU pivoting engine instead: an LP `min c'x s.t. l <= Ax <= u` is the GAVI with
M = [0 -A'], o = c (`LPSolver`).  Batches of LPs go to the device in one call.

Projection: the reference goes H-rep -> V-rep -> project -> H-rep through Polyhedra.jl's double
description (sets.jl:501-523).  Here the same set is obtained by eliminating the dropped
coordinates directly: through an equality row when one exists, else by Fourier-Motzkin
combination, with LP-based redundancy removal at the end.  The resulting H-rep can list different
(equivalent) rows than Polyhedra.jl would; everything downstream is geometric.
"""
import math

import numpy as np

from .model import INF, Poly

EQ_TOL = 1e-6          # sets.jl:420  isapprox(s.l, s.u; atol=tol) -> hyperplane


# --------------------------------------------------------------------------------------------
# membership (sets.jl:820-853) -- host form; the batched device form is Engine.halfspace_in
# --------------------------------------------------------------------------------------------
def contains(P, x, tol=1e-6, closed=False):
    ax = P.A @ np.asarray(x, dtype=float) if len(P) else np.zeros(0)
    for i in range(len(P)):
        lo_strict = bool(P.rl[i]) and not closed
        up_strict = bool(P.ru[i]) and not closed
        lo_ok = (P.l[i] - tol < ax[i]) if lo_strict else (P.l[i] - tol <= ax[i])
        up_ok = (ax[i] - tol < P.u[i]) if up_strict else (ax[i] - tol <= P.u[i])
        if not (lo_ok and up_ok):
            return False
    return True


def contains_prefix(P, x, engine, tol=1e-6):
    """sets.jl:826-847: `x in poly` when x gives only the first len(x) coordinates -- is there a completion
    y with  l - A_p x <= A_y y <= u - A_p x ?  The reference asks OSQP for the least-norm y (status 3 =
    primal infeasible -> false); here the same QP  min 0.5 |y|^2  is solved on the device engine."""
    x = np.asarray(x, dtype=float)
    n, d = len(x), P.dim
    if n == d:
        return contains(P, x, tol=tol)
    if len(P) == 0:
        return True
    from .qp import SolverFailure, solve_qp
    shift = P.A[:, :n] @ x
    try:
        y = solve_qp(engine, np.eye(d - n), np.zeros(d - n), P.A[:, n:], P.l - shift - tol, P.u - shift + tol)
    except SolverFailure:
        return False
    return bool(np.all(P.A[:, n:] @ y >= P.l - shift - 2 * tol) and np.all(P.A[:, n:] @ y <= P.u - shift + 2 * tol))


def intersect(*polys):
    """poly_intersect (sets.jl:936-968): the conjunction of all slices."""
    polys = [p for p in polys if p is not None]
    d = polys[0].dim
    A = np.vstack([p.A for p in polys]) if any(len(p) for p in polys) else np.zeros((0, d))
    cat = lambda f, dt: np.concatenate([np.asarray(getattr(p, f), dtype=dt) for p in polys]) if len(A) else np.zeros(0, dt)
    return Poly(A, cat("l", float), cat("u", float), cat("rl", bool), cat("ru", bool), normalize=False)


def complement(P):
    """sets.jl:918-930: one open half-space per finite bound."""
    out = []
    for i in range(len(P)):
        a = P.A[i:i + 1]
        if not math.isinf(P.l[i]):
            out.append(Poly(a, [-INF], [P.l[i]], [True], [not P.rl[i]], normalize=False))
        if not math.isinf(P.u[i]):
            out.append(Poly(a, [P.u[i]], [INF], [not P.ru[i]], [True], normalize=False))
    return out


def simplify(P, tol=1e-6):
    """sets.jl:255-305: merge slices with the same normal, keeping the tighter bounds."""
    m, d = len(P), P.dim
    if m == 0:
        return Poly(np.zeros((0, d)), [], [])
    A = P.A
    nonzero = np.sqrt(np.einsum("ij,ij->i", A, A)) > tol
    # all pairwise distances between normals at once (m is a few dozen); the sequential part below only looks up
    # "first kept slice with this normal" (isapprox(k, s.a; atol=tol))
    diff = A[:, None, :] - A[None, :, :]
    close = np.sqrt(np.einsum("ijk,ijk->ij", diff, diff)) <= tol
    L, U, RL, RU = P.l.tolist(), P.u.tolist(), P.rl.tolist(), P.ru.tolist()
    keep, kept_rows = [], []      # [row, l, u, rl, ru]
    for i in range(m):
        l, u, rl, ru = L[i], U[i], RL[i], RU[i]
        k = None
        if kept_rows:
            hit = np.flatnonzero(close[i, kept_rows])
            if len(hit):
                k = keep[int(hit[0])]
        if k is not None:
            if k[1] > l + tol:
                nl, nrl = k[1], k[3]
            elif l > k[1] + tol:
                nl, nrl = l, rl
            else:
                nl, nrl = 0.5 * (k[1] + l) if not (math.isinf(k[1]) and math.isinf(l)) else l, (True if k[3] else rl)
            if k[2] < u - tol:
                nu, nru = k[2], k[4]
            elif u < k[2] - tol:
                nu, nru = u, ru
            else:
                nu, nru = 0.5 * (k[2] + u) if not (math.isinf(k[2]) and math.isinf(u)) else u, (True if k[4] else ru)
            k[1], k[2], k[3], k[4] = nl, nu, nrl, nru
        elif nonzero[i]:
            keep.append([i, l, u, rl, ru]); kept_rows.append(i)
    if not keep:
        return Poly(np.zeros((0, d)), [], [])
    return Poly(A[kept_rows], [k[1] for k in keep], [k[2] for k in keep], [k[3] for k in keep], [k[4] for k in keep])


def poly_slice(P, fixed):
    """sets.jl:532-542: fix some coordinates (fixed: dict index -> value), drop them."""
    d = P.dim
    keep = [j for j in range(d) if j not in fixed]
    shift = np.zeros(len(P))
    for j, v in fixed.items():
        shift += P.A[:, j] * v
    return Poly(P.A[:, keep], P.l - shift, P.u - shift, P.rl, P.ru)


# --------------------------------------------------------------------------------------------
# LPs through the GPU engine
# --------------------------------------------------------------------------------------------
class LPSolver:
    """min c'x (+ 0.5 rho |x|^2 if rho) s.t. l <= Ax <= u, by the pivoting engine on the device.

    Returns dict(status, x, lam, obj); status 1 = solved, anything else = infeasible or unbounded
    (the reference reads OSQP's status_val the same coarse way: solved or not)."""

    def __init__(self, engine):
        self.engine = engine
        self.calls = 0
        # Geometric predicates memoised by polyhedron (Poly hashes / compares by its 5-digit slice keys, the
        # reference's own notion of equal sets, sets.jl:104-112,141-146): across the instances of a batch the
        # same pieces come back again and again.
        self.memo = {}
        from .batching import plain_once
        self.once = getattr(engine, "once", plain_once)      # a BatchingEngine memoises across its instance threads

    def solve(self, c, A, l, u, x0=None, rho=0.0):
        c = np.asarray(c, dtype=float)
        A = np.asarray(A, dtype=float).reshape(-1, len(c))
        n, m = len(c), A.shape[0]
        g = dict(M=np.hstack([rho * np.eye(n), -A.T]), N=np.zeros((n, 0)), o=c, l1=np.full(n, -INF), u1=np.full(n, INF),
                 A=np.hstack([A, np.zeros((m, m))]), B=np.zeros((m, 0)), l2=np.asarray(l, float), u2=np.asarray(u, float))
        z0 = np.zeros((1, n + m))
        if x0 is not None:
            z0[0, :n] = x0
        ret = self.engine.gavi_solve(g, np.zeros((1, 0)), z0)
        self.calls += 1
        x = ret["z"][0, :n]
        return dict(status=int(ret["status"][0]), x=x, lam=ret["z"][0, n:], obj=float(c @ x))


def exemplar(P, lp, tol=1e-2):
    """sets.jl:591-642.  Returns (empty, example)."""
    memo = getattr(lp, "memo", None)
    if memo is None:
        return _exemplar(P, lp, tol)
    return lp.once(memo, ("exemplar", P.exact_key, tol), lambda: _exemplar(P, lp, tol))


def _exemplar(P, lp, tol):
    n = len(P)
    if n == 0:
        return False, None
    d = P.dim
    open_low = P.rl & ~np.isinf(P.l)
    open_hi = P.ru & ~np.isinf(P.u)
    if (np.allclose(P.l, P.u, atol=tol, rtol=tol) and not open_low.any() and not open_hi.any() and n == d
            and not np.isinf(P.l).any()):
        try:
            x = np.linalg.solve(P.A, P.l)
            if np.allclose(P.A @ x, P.l, atol=tol, rtol=tol):
                return False, x
            return True, None
        except np.linalg.LinAlgError:
            pass
    # min eps  s.t.  A x + eps >= l,  -A x + eps >= -u
    AA = np.hstack([np.vstack([P.A, -P.A]), np.ones((2 * n, 1))])
    ll = np.concatenate([P.l, -P.u])
    fin = ~np.isinf(ll)                                 # rows with an infinite right-hand side constrain nothing
    c = np.zeros(d + 1); c[-1] = 1.0
    res = lp.solve(c, AA[fin], ll[fin], np.full(int(fin.sum()), INF))
    if res["status"] != 1:
        return False, None                              # unbounded below: a whole cone of interior points
    eps = res["x"][-1]
    x = res["x"][:-1]
    if eps > tol:
        return True, None
    if eps > -tol:
        y = np.zeros(2 * n); y[fin] = res["lam"]
        active_l, active_u = np.abs(y[:n]) > tol, np.abs(y[n:]) > tol
        if (active_l & open_low).any() or (active_u & open_hi).any():
            return True, None
    return False, x


def isempty(P, lp, tol=1e-4, x=None):
    """sets.jl:647-655: the cheap membership short-circuit, then the exemplar LP."""
    if x is not None and contains(P, x):
        return False
    return exemplar(P, lp, tol=tol)[0]


def issubset(P1, P2, lp, tol=1e-6):
    """sets.jl:377-407: for every finite bound of P2 minimise the bound's direction over P1."""
    memo = getattr(lp, "memo", None)
    if memo is None:
        return _issubset(P1, P2, lp, tol)
    return lp.once(memo, ("issubset", P1.exact_key, P2.exact_key, tol), lambda: _issubset(P1, P2, lp, tol))


def _issubset(P1, P2, lp, tol):
    for i in range(len(P2)):
        for bound, dirn in ((P2.l[i], 1.0), (P2.u[i], -1.0)):
            if math.isinf(bound):
                continue
            if len(P1) == 0:
                return False
            res = lp.solve(dirn * P2.A[i], P1.A, P1.l, P1.u)
            if res["status"] != 1:
                return False
            if res["obj"] < dirn * bound - tol:
                return False
    return True


def remove_subsets(polys, lp):
    """sets.jl:889-902."""
    k = len(polys)
    is_sub = [False] * k
    for i in range(k):
        if any(i != j and not is_sub[j] and issubset(polys[i], polys[j], lp) for j in range(k)):
            is_sub[i] = True
    return [p for p, s in zip(polys, is_sub) if not s]


# --------------------------------------------------------------------------------------------
# projection (replaces sets.jl:501-523)
# --------------------------------------------------------------------------------------------
def _split(P, tol=EQ_TOL):
    """Rows as equalities (a, b) and inequalities a'x <= b (get_Polyhedron_hrep, sets.jl:415-432)."""
    eq, ineq = [], []
    for i in range(len(P)):
        a, l, u = P.A[i], P.l[i], P.u[i]
        if not math.isinf(l) and not math.isinf(u) and abs(l - u) <= tol:
            eq.append((a.copy(), u))
        else:
            if not math.isinf(l):
                ineq.append((-a, -l))
            if not math.isinf(u):
                ineq.append((a.copy(), u))
    return eq, ineq


def _dedupe(rows, tol=1e-9):
    out, seen = [], set()
    for a, b in rows:
        nrm = np.abs(a).max() if len(a) else 0.0
        if nrm <= tol:
            continue                                    # 0 <= b rows (b >= 0 for a nonempty set) carry nothing
        a, b = a / nrm, b / nrm
        key = (tuple(np.round(a, 9) + 0.0), round(b, 9) + 0.0)
        if key in seen:
            continue
        seen.add(key)
        out.append((a, b))
    return out


def project(P, keep_dims, lp=None, tol=1e-9):
    """Projection of the closed polyhedron P onto the coordinates keep_dims (in that order)."""
    d = P.dim
    eq, ineq = _split(P)
    drop = [j for j in range(d) if j not in set(keep_dims)]
    for j in drop:
        piv = max(range(len(eq)), key=lambda k: abs(eq[k][0][j]), default=None)
        if piv is not None and abs(eq[piv][0][j]) > tol:
            a0, b0 = eq.pop(piv)
            sub = lambda a, b: (a - (a[j] / a0[j]) * a0, b - (a[j] / a0[j]) * b0)
            eq = [sub(a, b) for a, b in eq]
            ineq = [sub(a, b) for a, b in ineq]
        else:
            pos = [(a, b) for a, b in ineq if a[j] > tol]
            neg = [(a, b) for a, b in ineq if a[j] < -tol]
            zer = [(a, b) for a, b in ineq if abs(a[j]) <= tol]
            for ap, bp in pos:
                for an, bn in neg:
                    zer.append((ap / ap[j] - an / an[j], bp / ap[j] - bn / an[j]))
            ineq = zer
        for rows in (eq, ineq):
            for a, _ in rows:
                a[j] = 0.0
        eq, ineq = _dedupe(eq), _dedupe(ineq)
        if lp is not None and len(ineq) > 24:
            ineq = _irredundant(eq, ineq, lp)
    if lp is not None:
        ineq = _irredundant(eq, ineq, lp)
    kd = list(keep_dims)
    rows = [(a[kd], b, b) for a, b in eq] + [(a[kd], -INF, b) for a, b in ineq]
    if not rows:
        return Poly(np.zeros((0, len(kd))), [], [])
    return Poly(np.array([r[0] for r in rows]), [r[1] for r in rows], [r[2] for r in rows])


def _irredundant(eq, ineq, lp, tol=1e-7):
    """Drop every inequality implied by the others (max a'x over the rest <= b)."""
    keep = list(ineq)
    i = 0
    while i < len(keep):
        a, b = keep[i]
        others = keep[:i] + keep[i + 1:]
        A = np.array([r[0] for r in eq] + [r[0] for r in others]).reshape(-1, len(a))
        l = np.array([r[1] for r in eq] + [-INF] * len(others))
        u = np.array([r[1] for r in eq] + [r[1] for r in others])
        res = lp.solve(-a, A, l, u) if len(A) else dict(status=0)
        if res["status"] == 1 and -res["obj"] <= b + tol:
            keep.pop(i)
        else:
            i += 1
    return keep
