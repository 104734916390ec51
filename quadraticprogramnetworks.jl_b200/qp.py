"""Row A8 of SURVEY.md: `solve_qp` and `check_qp_convexity`
(/root/reference/src/qp_processing.jl:1-55) and `implicit_bounds`
(/root/reference/src/sets.jl:660-713).

Every QP / LP the reference hands to OSQP or PATH here goes through the device engine
(`qpn_gavi_solve_batched`); the 2 m bound LPs of `implicit_bounds` share one constraint matrix and
run as ONE batched launch (the cost vector rides as the GAVI parameter w).  The final step of the
convexity check -- a null-space basis and the eigenvalues of an nd x nd projected Hessian, once per
node at set-up time and off by default (`check_convexity=false`, programs.jl:74) -- is host linear
algebra, as in the reference (`svd`, `eigen`)."""
import numpy as np

from .model import INF


class SolverFailure(RuntimeError):
    """`error("Solver failure. ...")`, qp_processing.jl:8,30."""


def solve_qp(engine, Q, q, A, l, u, x0=None):
    """min 0.5 x'Qx + q'x  s.t.  l <= A x <= u  ->  x.  Both reference branches (:OSQP and :PATH,
    qp_processing.jl:2-33) are the same lifted KKT system here, solved by complementary pivoting."""
    Q, q = np.asarray(Q, float), np.asarray(q, float)
    n = len(q)
    A = np.asarray(A, float).reshape(-1, n)
    m = A.shape[0]
    g = dict(M=np.hstack([Q, -A.T]), N=np.zeros((n, 0)), o=q, l1=np.full(n, -INF), u1=np.full(n, INF),
             A=np.hstack([A, np.zeros((m, m))]), B=np.zeros((m, 0)), l2=np.asarray(l, float), u2=np.asarray(u, float))
    z0 = np.zeros((1, n + m))
    if x0 is not None:
        z0[0, :n] = x0
    ret = engine.gavi_solve(g, np.zeros((1, 0)), z0)
    if int(ret["status"][0]) != 1:
        raise SolverFailure(f"Solver failure. Status value is {int(ret['status'][0])}")
    return ret["z"][0, :n]


def implicit_bounds(engine, A, l, u, tol=1e-4):
    """sets.jl:660-713: rows whose lower and upper value over the polyhedron coincide.
    Returns (implicitly_equality (bool, m), vals (m))."""
    A = np.atleast_2d(np.asarray(A, float))
    l, u = np.asarray(l, float), np.asarray(u, float)
    m, d = A.shape
    eq = np.isclose(l, u, rtol=0.0, atol=tol)
    vals = np.where(eq, 0.5 * (l + u), INF)
    todo = np.flatnonzero(~eq)
    if len(todo) == 0:
        return eq, vals
    # min / max a_i'x over {l <= Ax <= u} for every remaining row: 2 |todo| LPs, one launch
    g = dict(M=np.hstack([np.zeros((d, d)), -A.T]), N=np.eye(d), o=np.zeros(d), l1=np.full(d, -INF), u1=np.full(d, INF),
             A=np.hstack([A, np.zeros((m, m))]), B=np.zeros((m, d)), l2=l, u2=u)
    W = np.vstack([A[todo], -A[todo]])
    ret = engine.gavi_solve(g, W, np.zeros((len(W), d + m)))
    k = len(todo)
    x = ret["z"][:, :d]
    ok = ret["status"] == 1
    # a failed LP is either unbounded in that direction or the set is empty; the start is projected onto the set first
    # (find_closest_feasible!), so an empty set shows as a point that violates the rows
    feas = np.all((x @ A.T >= l - 1e-6) & (x @ A.T <= u + 1e-6), axis=1)
    if ok.any() and not feas[ok].all():
        raise SolverFailure("Empty set")
    lo = np.where(ok[:k], np.einsum("ij,ij->i", A[todo], x[:k]), -INF)
    hi = np.where(ok[k:], np.einsum("ij,ij->i", A[todo], x[k:]), INF)
    same = np.isclose(lo, hi, rtol=0.0, atol=tol)
    eq[todo] = same
    vals[todo] = np.where(same, 0.5 * (lo + hi), INF)
    return eq, vals


def check_qp_convexity(engine, Q, A, l, u, dec_inds, pid=0, tol=1e-6):
    """qp_processing.jl:37-54: Q restricted to the decision variables must be positive semidefinite on the
    null space of the implicitly-equality rows.  Raises like the reference; returns the smallest eigenvalue."""
    Q = np.asarray(Q, float)
    A = np.atleast_2d(np.asarray(A, float))
    dec = list(dec_inds)
    if A.shape[0]:
        eq, _ = implicit_bounds(engine, A, l, u, tol=tol)
        Ae = A[np.ix_(np.flatnonzero(eq), dec)]
    else:
        Ae = np.zeros((0, len(dec)))
    if Ae.shape[0]:
        _, S, Vt = np.linalg.svd(Ae, full_matrices=True)
        r = int(np.linalg.matrix_rank(np.diag(S))) if len(S) else 0
        Z = Vt.T[:, r:]
    else:
        Z = np.eye(len(dec))
    if Z.shape[1] == 0:
        return INF
    QQ = Z.T @ Q[np.ix_(dec, dec)] @ Z
    lam_min = float(np.linalg.eigvalsh(QQ + QQ.T).min())
    if not lam_min > -tol:
        raise ValueError(f"QP {pid} is not convex. Exiting.")
    return lam_min
