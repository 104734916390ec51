"""CPU suite (-m "not gpu"): pins the oracle.

The reference's AVI arithmetic lives in closed-source PATH and Julia is not installed, so
the oracle is pinned on: KAT-0, the reference's closed forms for
simple_bilevel's lower level, check_avi_solution residuals, optimal values from scipy's LP /
QP solvers, uniqueness for the strongly monotone four-player game, agreement of the Python
statement with the C port bit for bit, and the committed golden fixtures.
"""
import json
import math
import os

import numpy as np
import pytest
from scipy.optimize import linprog, minimize

from oracle import avi_pivot, cport, examples, qpn_ref
from tests import problems

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INF = math.inf


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def _flt(v):
    return [float(x) for x in v]


# ---- KAT-0: assembly + AVI solve of simple_bilevel's lower level ------------------------
def test_kat0_assembly_matches_hand_derivation():
    kat = load("simple_bilevel_kat.json")["kat0"]
    net = examples.simple_bilevel()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[2], {})
    avi = qpn_ref.convert(g)
    assert np.array_equal(avi["M"], np.array(kat["M"], dtype=float))
    assert np.array_equal(avi["N"], np.array(kat["N"], dtype=float))
    assert np.array_equal(avi["l"], np.array(_flt(kat["l"]))) and np.array_equal(avi["u"], np.array(_flt(kat["u"])))
    assert dec == [3] and par == [0, 1, 2]


@pytest.mark.parametrize("solver", ["python", "c"])
def test_kat0_solution_closed_form(solver):
    net = examples.simple_bilevel()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[2], {})
    rng = np.random.default_rng(0)
    for xv in [-3.0, -1.0, -1e-3, 0.0, 1e-3, 0.7, 2.0] + list(rng.normal(size=10)):
        for y0 in (0.0, 1.5, -0.5):
            x = np.array([0.3, -0.2, xv, y0])
            z0 = np.array([x[3], 0.0, 0.0])
            ret = qpn_ref.solve_gavi(g, z0, x[par]) if solver == "python" else cport.gavi_solve(g, z0, x[par])
            assert ret["status"] == 1
            expect = [max(xv, 0.0), 0.0, max(-2 * xv, 0.0), max(xv, 0.0)]
            assert np.allclose(ret["z_full"], expect, atol=1e-12), (xv, y0, ret["z_full"])


# ---- Python statement == C port, bit for bit ----------------------------------------------
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_c_port_equals_python_statement(kind):
    rng = np.random.default_rng(7 + kind)
    for _ in range(25):
        Q, c, A, l, u, z0 = problems.random_qp(rng, kind)
        g = problems.qp_gavi(Q, c, A, l, u)
        r1 = qpn_ref.solve_gavi(g, z0.copy(), np.zeros(0))
        r2 = cport.gavi_solve(g, z0.copy(), np.zeros(0))
        assert r1["status"] == r2["status"] and r1["pivots"] == r2["pivots"]
        assert np.array_equal(r1["basis"], r2["basis"]) and np.array_equal(r1["z_full"], r2["z_full"])


def test_c_port_equals_python_on_examples():
    rng = np.random.default_rng(3)
    for net, lev in [(examples.robust_avoid_simple(), 3), (examples.four_player_matrix_game(), 1), (examples.simple_bilevel(), 2)]:
        g, dec, par = qpn_ref.level_gavi(net, net.depth[lev], {})
        for _ in range(4):
            x = net.default_init + rng.normal(size=net.n_vars)
            w = x[par]
            z0 = np.concatenate([x[dec], np.zeros(g["M"].shape[1] - len(dec))])
            r1 = qpn_ref.solve_gavi(g, z0.copy(), w)
            r2 = cport.gavi_solve(g, z0.copy(), w)
            assert r1["status"] == r2["status"] == 1 and r1["pivots"] == r2["pivots"]
            assert np.array_equal(r1["z_full"], r2["z_full"]) and np.array_equal(r1["basis"], r2["basis"])
            assert np.array_equal(qpn_ref.comp_indices(g, r1["z"], w), cport.comp_indices(g, r2["z"], w))


# ---- optimality against independent solvers --------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_optimal_value_matches_scipy(kind):
    rng = np.random.default_rng(40 + kind)
    for _ in range(60):
        Q, c, A, l, u, z0 = problems.random_qp(rng, kind)
        n = len(c)
        g = problems.qp_gavi(Q, c, A, l, u)
        ret = cport.gavi_solve(g, z0, np.zeros(0))
        assert ret["status"] == 1
        x = ret["z"][:n]
        obj = 0.5 * x @ Q @ x + c @ x
        fin = np.isfinite(u)
        if kind == 0:
            ref = linprog(c, A_ub=np.vstack([-A, A[fin]]), b_ub=np.concatenate([-l, u[fin]]), bounds=[(None, None)] * n, method="highs").fun
        else:
            cons = [{"type": "ineq", "fun": lambda v, A=A, l=l: A @ v - l}]
            if fin.any():
                cons.append({"type": "ineq", "fun": lambda v, A=A, u=u, fin=fin: u[fin] - A[fin] @ v})
            x0 = np.linalg.lstsq(A, np.where(np.isfinite(u), 0.5 * (l + np.where(fin, u, l)), l + 0.5), rcond=None)[0]
            ref = minimize(lambda v: 0.5 * v @ Q @ v + c @ v, x0, jac=lambda v: Q @ v + c, constraints=cons, method="SLSQP",
                           options=dict(ftol=1e-13, maxiter=800)).fun
        assert abs(obj - ref) <= 2e-6 * (1 + abs(ref)), (kind, obj, ref)
        assert np.all(A @ x >= l - 1e-7) and np.all(A @ x <= u + 1e-7)


def test_four_player_unique_equilibrium_and_residual():
    net, g, avi, dec, par = problems.fp_avi()
    rng = np.random.default_rng(5)
    X, z0 = problems.fp_starts(rng, 64)
    z, st, pv, bs = cport.avi_solve_batched(avi["M"], np.tile(avi["o"], (64, 1)), avi["l"], avi["u"], z0)
    assert (st == 1).all()
    # interior equilibrium = solution of the stacked first-order conditions  J x = -q
    J = np.vstack([net.qps[p]["Q"][net.decision_inds(p), :] for p in net.depth[1]])
    q = np.concatenate([net.qps[p]["q"][net.decision_inds(p)] for p in net.depth[1]])
    xs = np.linalg.solve(J, -q)
    assert np.all(np.abs(xs) < 5)
    assert np.allclose(z[:, :8], xs, atol=1e-10)
    for k in range(64):
        bad, cnt, r = cport.check_avi(avi["M"], avi["o"], avi["l"], avi["u"], z[k])
        assert not bad and np.abs(r).max() < 1e-9


def test_robust_avoid_bottom_level_matches_linprog():
    rng = np.random.default_rng(6)
    net, X = problems.ra_inits(rng, 24)
    for x in X:
        xo, ret = qpn_ref.solve_qep(net, net.depth[3], x, {}, solver=lambda *a: cport.avi_solve(*a))
        assert ret["status"] == 1
        for pid in net.depth[3]:
            qp, c = net.qps[pid], net.cons[net.qps[pid]["cons"][0]]
            dv = net.decision_inds(pid)
            pr = [i for i in range(net.n_vars) if i not in dv]
            res = linprog(qp["q"][dv], A_ub=-c["A"][:, dv], b_ub=-(c["l"] - c["A"][:, pr] @ x[pr]), bounds=[(None, None)] * 3, method="highs")
            assert abs(res.fun - xo[dv][2]) < 1e-8


# ---- failure modes --------------------------------------------------------------------------
def test_unbounded_and_infeasible_are_not_success():
    # min -x s.t. x >= 0: unbounded
    g = problems.qp_gavi(np.zeros((1, 1)), np.array([-1.0]), np.array([[1.0]]), np.array([0.0]), np.array([INF]))
    assert cport.gavi_solve(g, np.array([1.0, 0.0]), np.zeros(0))["status"] != 1
    # x <= -1 and x >= 1: infeasible
    g = problems.qp_gavi(np.eye(1), np.array([0.0]), np.array([[1.0], [1.0]]), np.array([-INF, 1.0]), np.array([-1.0, INF]))
    assert cport.gavi_solve(g, np.array([0.0, 0.0, 0.0]), np.zeros(0))["status"] != 1


# ---- check_avi_solution (avi.jl:148-156) ------------------------------------------------------
def test_check_avi_counts():
    M = np.array([[2.0, 0.0], [0.0, 1.0]])
    q = np.array([-2.0, 1.0])
    l, u = np.array([0.0, 0.0]), np.array([5.0, 5.0])
    assert cport.check_avi(M, q, l, u, np.array([1.0, 0.0]))[1] == 0          # r = [0, 1], z2 at lower
    assert cport.check_avi(M, q, l, u, np.array([1.0, 0.5]))[1] == 1          # r2 > 0 but z2 not at lower
    assert cport.check_avi(M, q, l, u, np.array([-0.1, 0.0]))[1] == 2         # below lower and r1 < 0 off upper
    assert cport.check_avi(M, q, l, u, np.array([5.0 + 1e-3, 0.0]))[1] == 2   # above upper, r1 > 0 off lower


# ---- Base.in (sets.jl:820-853) ------------------------------------------------------------------
def test_halfspace_in_relations_and_tolerance():
    A = np.array([[1.0, 0.0]])
    for (x1, tol, rl, expect) in [(0.0, 1e-6, 0, True), (0.0, 1e-6, 1, True), (-1e-6, 1e-6, 0, True), (-1e-6, 1e-6, 1, False),
                                  (-2e-6, 1e-6, 0, False), (-5e-4, 1e-3, 0, True)]:
        got = cport.halfspace_in(A, [0.0], [INF], np.array([x1, 0.0]), tol, rl=[rl], ru=[0])
        assert got == expect == qpn_ref.in_slice([x1, 0.0], A[0], 0.0, INF, bool(rl), False, tol), (x1, tol, rl)
    # Slice normalisation flips rows with a negative leading coefficient (sets.jl:82-87)
    P = qpn_ref.Poly([[-2.0, 4.0]], [-INF], [6.0])
    assert np.allclose(P.A, [[1.0, -2.0]]) and P.l[0] == -3.0 and P.u[0] == INF


# ---- comp_indices (avi_solutions.jl:511-612) --------------------------------------------------------
def test_comp_indices_masks():
    l = np.array([0.0, 0.0, 0.0, 1.0, -INF])
    u = np.array([1.0, 1.0, 1.0, 1.0, INF])
    z = np.array([0.0, 0.5, 1.0, 1.0, 3.0])
    r = np.array([2.0, 0.0, -2.0, 7.0, 0.0])
    m = qpn_ref.comp_indices_block(l, u, r, z)
    assert list(m) == [1, 2, 4, 8, 2]
    # weakly active: z at lower with r = 0 is in sets 1 and 2
    assert qpn_ref.comp_indices_block(np.array([0.0]), np.array([1.0]), np.array([0.0]), np.array([0.0]))[0] == 3


# ---- verify_solution (qp_processing.jl:57-149) ---------------------------------------------------------
def test_verify_solution_simple_bilevel_lower_node():
    net = examples.simple_bilevel()
    view = qpn_ref.node_view(net, 1)
    for x, expect_sol, expect_how in [([0, 0, 1.0, 1.0], True, 2), ([0, 0, -1.0, 0.0], True, 2), ([0, 0, 1.0, 0.5], False, 4),
                                      ([0, 0, 1.0, -0.5], False, 0), ([0, 0, 0.0, 0.0], True, 2)]:
        sol, lam, how, act = cport.verify_solution(*view, np.array(x, dtype=float))
        assert sol == expect_sol and how == expect_how, (x, sol, how)
    sol, lam, how, act = cport.verify_solution(*view, np.array([0, 0, -1.0, 0.0]))
    assert np.allclose(lam, [2.0]) and list(act) == [1]            # 2(y - x) = lambda, y >= 0 active


def test_verify_solution_fallback_branch_is_reached():
    """Three active constraints in the plane: the basic least-squares solution has a negative
    multiplier, the sign-constrained fallback (qp_processing.jl:129-146) finds a valid one."""
    nv = 2
    Qd, qd = np.zeros((2, 2)), np.array([1.0, 1.0])                 # gradient (1, 1)
    A = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, -1.0]])            # x >= 0, y >= 0, x - y >= 0 all active at 0
    l, u = np.zeros(3), np.full(3, INF)
    hows = set()
    for perm in ([0, 1, 2], [2, 0, 1], [2, 1, 0], [1, 2, 0]):
        sol, lam, how, act = cport.verify_solution(Qd, qd, A[perm], l, u, np.array([0, 1], np.int32), np.zeros(nv))
        assert sol and np.allclose(A[perm].T @ lam, qd, atol=1e-8) and (lam > -1e-4).all()
        hows.add(how)
    assert hows <= {2, 3}


# ---- level loop --------------------------------------------------------------------------------------------
def test_level_loop_c_equals_python_and_goldens():
    gold = load("oracle_robust_avoid_bottom_level.json")
    net = examples.robust_avoid_simple()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    L = cport.Level(net.n_vars, [qpn_ref.node_view(net, p) for p in net.depth[3]], g, dec, par, 150, None)
    X = np.array(gold["inits"])
    r = L.solve(X, threads=2)
    assert np.array_equal(r["x"], np.array(gold["x"])) and list(r["pivots"]) == gold["pivots"] and list(r["iters"]) == gold["iters"]
    assert np.array_equal(r["lam"], np.array(gold["lam"]))
    for k in range(len(X)):
        ro = qpn_ref.solve_level_bottom(net, 3, X[k], None)
        assert ro["solved"] and ro["iters"] == r["iters"][k] and ro["pivots"] == r["pivots"][k] and np.array_equal(ro["x"], r["x"][k])


def test_four_player_avi_goldens():
    gold = load("oracle_four_player_avi.json")
    net, g, avi, dec, par = problems.fp_avi()
    X = np.array(gold["inits"])
    z0 = np.zeros((len(X), 32)); z0[:, :8] = X; z0[:, 24:] = X
    z, st, pv, bs = cport.avi_solve_batched(avi["M"], np.tile(avi["o"], (len(X), 1)), avi["l"], avi["u"], z0, threads=2)
    assert np.array_equal(z, np.array(gold["z"])) and list(st) == gold["status"] and list(pv) == gold["pivots"]
    assert np.array_equal(bs, np.array(gold["basis"], dtype=np.int8))


def test_cycle_check_semantics():
    """algorithm.jl:24: proj_vals ~ previous (isapprox, rtol = sqrt(eps)) ends the solve unsolved."""
    assert qpn_ref.projections_equal([1.0, 2.0], [1.0, 2.0 + 1e-9])
    assert not qpn_ref.projections_equal([1.0, 2.0], [1.0, 2.0 + 1e-6])


# ---- frozen rows (oracle/avi_pivot.py: freeze) ---------------------------------------------------------------
def test_frozen_rows_change_no_decision_and_only_last_bits():
    """The rows of free basics are frozen after phase 0 and their variables evaluated once at the end.  Against the
    same solve with the freeze switched off (every row carried through every pivot): same status, pivot count and
    bases; z equal to rounding; and the frozen components satisfy their own equations (M z + q)_k = 0 at least as
    well as the incremental ones."""
    from oracle import avi_pivot

    class Unfrozen(avi_pivot._Tab):
        def freeze(self):
            super().freeze()
            self.frozen, self.T0, self.beta0 = [], self.T0[:0], self.beta0[:0]

    def run(cls, M, q, l, u, z0):
        tab = cls(M, q, l, u, z0)
        tab.crash(); tab.repair()
        st = tab.lemke(50 * len(q) + 100)
        return st, tab.pivots, tab.basis_codes(), tab.solution(), len(tab.frozen)

    rng = np.random.default_rng(5)
    frozen_seen = 0
    for kind in (0, 1, 2):
        for _ in range(12):
            Q, c, A, lo, up, z0 = problems.random_qp(rng, kind)
            a = qpn_ref.convert(problems.qp_gavi(Q, c, A, lo, up))
            M, q, l, u = a["M"], a["o"], a["l"], a["u"]
            z0f = np.concatenate([z0, A @ z0[:len(c)]])
            s1, p1, b1, z1, nf = run(avi_pivot._Tab, M, q, l, u, z0f)
            s2, p2, b2, z2, _ = run(Unfrozen, M, q, l, u, z0f)
            frozen_seen += nf
            assert (s1, p1) == (s2, p2) and np.array_equal(b1, b2)
            assert np.allclose(z1, z2, rtol=1e-9, atol=1e-9)
            if s1 == 1:
                free = np.isinf(l) & np.isinf(u)
                r1, r2 = np.abs((M @ z1 + q)[free]).max(), np.abs((M @ z2 + q)[free]).max()
                assert r1 <= 10 * r2 + 1e-10
    assert frozen_seen > 100


def test_degenerate_lps_do_not_cycle():
    """Tie rule of the ratio test (t first, then largest |d|, then lowest row): highly degenerate LPs -- cones through
    the origin cut by a box, small integer data, so most ratio tests tie -- must all end in SUCCESS well inside the
    pivot cap, at HiGHS's optimal value.  (north_star names lexicographic tie-breaking; the rule here is deterministic
    but carries no anti-cycling proof, so this is the evidence: 300 degenerate instances, none runs into the cap.)"""
    from scipy.optimize import linprog
    rng = np.random.default_rng(77)
    worst = 0.0
    for trial in range(300):
        n = int(rng.integers(2, 6)); m = int(rng.integers(n + 1, 3 * n + 2))
        A = rng.integers(-1, 2, (m, n)).astype(float)
        A = A[np.abs(A).sum(1) > 0]
        m = len(A)
        c = rng.integers(-2, 3, n).astype(float)
        # A x >= 0 (every row through the origin) and -1 <= x <= 1: the origin is a vertex where all m rows are active
        AA = np.vstack([A, np.eye(n)]); l = np.concatenate([np.zeros(m), -np.ones(n)]); u = np.concatenate([np.full(m, INF), np.ones(n)])
        g = problems.qp_gavi(np.zeros((n, n)), c, AA, l, u)
        ret = cport.gavi_solve(g, np.zeros(n + len(l)), np.zeros(0))
        ref = linprog(c, A_ub=-A, b_ub=np.zeros(m), bounds=[(-1, 1)] * n, method="highs")
        assert ref.status == 0
        npiv_cap = 50 * (n + 2 * len(l)) + 100
        assert ret["status"] == 1 and ret["pivots"] < npiv_cap, (trial, ret["status"], ret["pivots"])
        worst = max(worst, ret["pivots"] / npiv_cap)
        assert abs(float(c @ ret["z"][:n]) - ref.fun) < 1e-7, (trial, float(c @ ret["z"][:n]), ref.fun)
    assert worst < 0.2


def test_julia_goldens_are_consumed_when_present():
    """tests/golden/julia/ is where output of julia/ref_julia.jl goes (true reference goldens: PATH's own solves).
    None exist in this image -- the test then only checks the schema file; any that are dropped in later are
    checked against the schema and solved on the oracle stand-in with the tolerances of
    scripts/compare_with_julia_goldens.py."""
    import json
    schema = json.load(open(os.path.join(GOLDEN, "julia_goldens.schema.json")))
    assert set(schema["required"]) == {"example", "threads", "seconds", "equilibria_per_s", "inits", "solved", "x"}
    jdir = os.path.join(GOLDEN, "julia")
    files = [f for f in os.listdir(jdir) if f.endswith("_goldens.json")] if os.path.isdir(jdir) else []
    for f in files:
        G = json.load(open(os.path.join(jdir, f)))
        assert set(schema["required"]) <= set(G), f
        assert len(G["inits"]) == len(G["solved"]) == len(G["x"]) and all(len(a) == len(b) for a, b in zip(G["inits"], G["x"])), f
        model = os.path.join(jdir, f.replace("_goldens.json", "_model.json"))
        assert os.path.exists(model), f"{f}: the model exported by julia/QPNCuda.jl: export_qpnet must sit next to the goldens"
        import qpn_b200
        from tests.native_oracle import oracle_net
        net = qpn_b200.load_net(model)
        X = np.array(G["inits"], dtype=float)
        if net.num_levels() == 1 and not net.options.gen_solution_map:
            continue                                                # flat games: compared on the device (scripts/compare_with_julia_goldens.py)
        ret = oracle_net(net, threads=2).solve_arrays(X)
        ref_solved = np.array(G["solved"], bool)
        both = ret["solved"] & ref_solved
        assert np.mean(ret["solved"] == ref_solved) > 0.9, f
        if G["example"] == "simple_bilevel":
            assert np.abs(ret["x"][both] - np.array(G["x"], dtype=float)[both]).max(initial=0.0) < 1e-4, f
