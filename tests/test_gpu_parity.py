"""GPU parity tests: every C-ABI entry point of libqpn_cuda against the C oracle on the
same seeded inputs.  Bar: status / pivots / basis / masks bit-exact; z within 1e-8
relative (north_star) -- and in fact bit-identical, which is asserted where noted."""
import numpy as np
import pytest

from oracle import cport, examples, qpn_ref
from tests import problems

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def assert_close(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    err = np.abs(a - b) / scale
    assert err.max(initial=0.0) <= RTOL, f"{what}: max rel err {err.max()}"


def test_avi_four_player_dense_and_csc(engine):
    rng = np.random.default_rng(11)
    net, g, avi, dec, par = problems.fp_avi()
    B = 512
    X, z0 = problems.fp_starts(rng, B)
    q = np.tile(avi["o"], (B, 1))
    zo, so, po, bo = cport.avi_solve_batched(avi["M"], q, avi["l"], avi["u"], z0)
    z, s, p, b = engine.avi_solve(avi["M"], q, avi["l"], avi["u"], z0)
    assert (so == 1).all()
    assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo)
    assert np.array_equal(z, zo), "z must match the oracle bit for bit"
    zc, sc, pc, bc = engine.avi_solve(None, q, avi["l"], avi["u"], z0, csc=problems.dense_to_csc(avi["M"], base=1), index_base=1)
    assert np.array_equal(zc, zo) and np.array_equal(sc, so) and np.array_equal(pc, po) and np.array_equal(bc, bo)
    # uniqueness (strongly monotone game): every start reaches the same equilibrium
    assert np.ptp(z[:, :8], axis=0).max() < 1e-9


def test_avi_tight_box_active_bounds(engine):
    """Same game in a box so small that bounds are active at the solution."""
    rng = np.random.default_rng(12)
    net, g, avi, dec, par = problems.fp_avi()
    l, u = avi["l"].copy(), avi["u"].copy()
    l[24:] = -0.4; u[24:] = 0.4
    B = 256
    X = rng.uniform(-0.4, 0.4, (B, 8))
    X[: B // 2] = rng.uniform(-2, 2, (B // 2, 8))            # half the starts are outside the box
    z0 = np.zeros((B, 32)); z0[:, :8] = X; z0[:, 24:] = X
    q = np.tile(avi["o"], (B, 1))
    zo, so, po, bo = cport.avi_solve_batched(avi["M"], q, l, u, z0)
    z, s, p, b = engine.avi_solve(avi["M"], q, l, u, z0)
    assert (so == 1).all() and (bo[:, 24:] != 2).any()
    assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo) and np.array_equal(z, zo)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_avi_random_qps_per_instance_matrix(engine, kind):
    rng = np.random.default_rng(100 + kind)
    for trial in range(12):
        n, m = int(rng.integers(2, 6)), None
        probs = []
        Q, c, A, l, u, z0 = problems.random_qp(rng, kind, n=n)
        m = len(l)
        B = 16
        Ms, qs, ls, us, z0s = [], [], [], [], []
        for _ in range(B):
            Q, c, A, l, u, z0 = problems.random_qp(rng, kind, n=n, m=m - (n if kind != 1 else 0))
            g = problems.qp_gavi(Q, c, A, l, u)
            avi = qpn_ref.convert(g)
            s0 = g["A"] @ z0
            Ms.append(avi["M"]); qs.append(avi["o"]); ls.append(avi["l"]); us.append(avi["u"]); z0s.append(np.concatenate([z0, s0]))
        Ms, qs, ls, us, z0s = map(np.array, (Ms, qs, ls, us, z0s))
        zo, so, po, bo = cport.avi_solve_batched(Ms, qs, ls, us, z0s)
        z, s, p, b = engine.avi_solve(Ms, qs, ls, us, z0s)
        assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo)
        assert_close(z, zo, "z")
        assert np.array_equal(z, zo)


def test_check_avi(engine):
    rng = np.random.default_rng(13)
    net, g, avi, dec, par = problems.fp_avi()
    B = 64
    X, z0 = problems.fp_starts(rng, B)
    q = np.tile(avi["o"], (B, 1))
    z, s, p, b = engine.avi_solve(avi["M"], q, avi["l"], avi["u"], z0)
    zz = np.vstack([z[:32], z0[:32]])                    # solutions and non-solutions
    bad, r = engine.check_avi(avi["M"], q[:64], avi["l"], avi["u"], zz)
    for k in range(64):
        sol_bad, cnt, rr = cport.check_avi(avi["M"], q[k], avi["l"], avi["u"], zz[k])
        assert cnt == bad[k] and np.array_equal(rr, r[k])
    assert (bad[:32] == 0).all() and (bad[32:] > 0).all()


def _gavi_cases():
    rng = np.random.default_rng(14)
    cases = []
    net = examples.simple_bilevel()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[2], {})
    X = rng.normal(size=(64, 4)) * 2
    cases.append(("simple_bilevel L2", g, dec, par, X))
    net, g, avi, dec, par = problems.fp_avi()
    cases.append(("four_player L1", g, dec, par, rng.uniform(-7, 7, (128, 8))))     # some outside the box -> presolve
    net, X = problems.ra_inits(rng, 128)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    cases.append(("robust_avoid L3", g, dec, par, X))
    return cases


def test_gavi_solve_and_comp_indices(engine):
    for name, g, dec, par, X in _gavi_cases():
        B = len(X)
        dz = g["M"].shape[1]
        w = X[:, par]
        z0 = np.zeros((B, dz)); z0[:, :len(dec)] = X[:, dec]
        ret = engine.gavi_solve(g, w, z0)
        masks = engine.comp_indices(g, ret["z"], w)
        presolved = 0
        for k in range(B):
            ro = cport.gavi_solve(g, z0[k], w[k])
            presolved += int(not np.array_equal(ro["z0_projected"], z0[k]))
            assert ro["status"] == ret["status"][k], name
            assert ro["pivots"] == ret["pivots"][k], name
            assert np.array_equal(ro["basis"], ret["basis"][k]), name
            assert np.array_equal(ro["z_full"], ret["z_full"][k]), name
            assert np.array_equal(cport.comp_indices(g, ro["z"], w[k]), masks[k]), name
        assert (ret["status"] == 1).all(), name
        if name != "simple_bilevel L2":
            assert presolved > 0, f"{name}: the presolve projection was never exercised"


def test_halfspace_in(engine):
    rng = np.random.default_rng(15)
    d = 6
    polys = []
    for _ in range(9):
        m = int(rng.integers(1, 40))
        A = rng.normal(size=(m, d)) * (rng.uniform(size=(m, d)) < 0.6)
        l = -rng.uniform(0.2, 2.5, m); u = rng.uniform(0.2, 2.5, m)
        l[rng.uniform(size=m) < 0.2] = -np.inf; u[rng.uniform(size=m) < 0.2] = np.inf
        rl = (rng.uniform(size=m) < 0.3).astype(np.uint8); ru = (rng.uniform(size=m) < 0.3).astype(np.uint8)
        polys.append((A, l, u, rl, ru))
    x = rng.normal(size=(200, d)) * 0.35
    for tol in (1e-6, 1e-3):
        got = engine.halfspace_in(polys, x, tol=tol)
        for j in range(len(x)):
            for p, (A, l, u, rl, ru) in enumerate(polys):
                assert got[j, p] == cport.halfspace_in(A, l, u, x[j], tol, rl, ru)
    assert got.any() and not got.all()


def _verify_cases():
    rng = np.random.default_rng(16)
    cases = []
    net = examples.four_player_matrix_game()
    r = qpn_ref.solve_level_bottom(net, 1, np.zeros(8))
    X = np.vstack([np.tile(r["x"], (8, 1)) + 1e-6 * rng.normal(size=(8, 8)), rng.uniform(-5, 5, (24, 8)),
                   np.clip(rng.uniform(-9, 9, (16, 8)), -5, 5)])
    for pid in net.depth[1]:
        cases.append((f"four_player node {pid}", qpn_ref.node_view(net, pid), X))
    net, Xr = problems.ra_inits(rng, 24)
    sols = np.array([qpn_ref.solve_level_bottom(net, 3, x)["x"] for x in Xr])
    Xall = np.vstack([sols, Xr])
    for pid in net.depth[3]:
        cases.append((f"robust_avoid node {pid}", qpn_ref.node_view(net, pid), Xall))
    net = examples.simple_bilevel()
    cases.append(("simple_bilevel node 1", qpn_ref.node_view(net, 1), np.array([[0, 0, 1.0, 1.0], [0, 0, -1.0, 0.0], [0, 0, 0.0, 0.0], [0, 0, 1.0, 0.5]])))
    return cases


def test_verify_solution(engine):
    seen = set()
    for name, view, X in _verify_cases():
        sol, lam, how, act = engine.verify_solution(view, X)
        for k in range(len(X)):
            so, lo, ho, ao = cport.verify_solution(*view, X[k])
            assert so == sol[k] and ho == how[k], (name, k, ho, how[k])
            assert np.array_equal(ao, act[k]), (name, k)
            assert np.array_equal(lo, lam[k]), (name, k, lo, lam[k])
            seen.add(int(ho))
    assert {0, 2, 4} <= seen, f"verify_solution branches exercised: {seen}"


def test_level_equilibrium_four_player(engine):
    import qpn_b200
    rng = np.random.default_rng(17)
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    proj = rng.normal(size=(4, 8))
    lv = qpn_b200.LevelArrays(8, [qpn_ref.node_view(net, p) for p in net.depth[1]], g, dec, par, max_iters=150, proj=proj)
    B = 256
    X = rng.uniform(-5, 5, (B, 8))
    ret = engine.level_equilibrium(lv, X)
    assert ret["solved"].all()
    for k in range(0, B, 8):
        ro = qpn_ref.solve_level_bottom(net, 1, X[k], proj)
        assert ro["solved"] and ro["iters"] == ret["iters"][k] and ro["pivots"] == ret["pivots"][k]
        assert np.array_equal(ro["x"], ret["x"][k])
        assert np.array_equal(ro["lam"], ret["lam"][k])
    assert np.ptp(ret["x"], axis=0).max() < 1e-9


def test_level_equilibrium_robust_avoid_bottom(engine):
    import qpn_b200
    rng = np.random.default_rng(18)
    net, X = problems.ra_inits(rng, 64)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    lv = qpn_b200.LevelArrays(net.n_vars, [qpn_ref.node_view(net, p) for p in net.depth[3]], g, dec, par, max_iters=150, proj=None)
    ret = engine.level_equilibrium(lv, X)
    for k in range(len(X)):
        ro = qpn_ref.solve_level_bottom(net, 3, X[k], None)
        assert ro["solved"] == ret["solved"][k] and ro["iters"] == ret["iters"][k] and ro["pivots"] == ret["pivots"][k], k
        assert np.array_equal(ro["x"], ret["x"][k])
        if ro["solved"]:
            assert np.array_equal(ro["lam"], ret["lam"][k])
    assert ret["solved"].all()


def test_resident_level_with_plans_matches_oracle(engine):
    """qpn_level_upload precomputes the instance-independent crash prefix (plans); results must
    still equal the oracle, which runs every pivot per instance."""
    import qpn_b200
    rng = np.random.default_rng(19)
    # four_player: inside the box (no presolve) and outside (presolve plan exercised)
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    proj = rng.normal(size=(4, 8))
    views = [qpn_ref.node_view(net, p) for p in net.depth[1]]
    lv = qpn_b200.ResidentLevel(engine, qpn_b200.LevelArrays(8, views, g, dec, par, max_iters=150, proj=proj))
    X = np.vstack([rng.uniform(-5, 5, (96, 8)), rng.uniform(-8, 8, (96, 8))])
    ret = lv.solve(X)
    L = cport.Level(8, views, g, dec, par, 150, proj)
    ro = L.solve(X)
    assert ro["solved"].all() and (ro["pivots"] != ro["pivots"][0]).any()
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    assert np.array_equal(ret["solved"].astype(bool), ro["solved"])
    lv.release()
    # robust_avoid bottom level: LP-like nodes, presolve on every instance
    net, X = problems.ra_inits(rng, 160)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    views = [qpn_ref.node_view(net, p) for p in net.depth[3]]
    lv = qpn_b200.ResidentLevel(engine, qpn_b200.LevelArrays(net.n_vars, views, g, dec, par, max_iters=150, proj=None))
    ret = lv.solve(X)
    ro = cport.Level(net.n_vars, views, g, dec, par, 150, None).solve(X)
    assert ro["solved"].all()
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    lv.release()


def test_simple_bilevel_known_answers_on_gpu(engine):
    """The reference's own test (test/simple_bilevel.jl:4-21) end to end on the device engine:
    every AVI solve, dual recovery, active-set classification and LP of the run is a kernel launch."""
    from tests.test_multilevel_cpu import check_simple_bilevel_kats
    before = engine.launches
    check_simple_bilevel_kats(engine)
    assert engine.launches > before + 50


def test_set_algebra_on_device_matches_oracle(engine):
    """SURVEY A11 / A12 / A13 directly on the device engine: exemplar / isempty, issubset / remove_subsets, complement,
    intersection emptiness, projection -- the hand cases of the CPU suite, then every predicate on a family of random
    polytopes with the same answers as with the oracle as the LP engine."""
    from tests.oracle_engine import OracleEngine
    from tests.test_multilevel_cpu import check_polyhedra_operations, set_algebra_answers
    before = engine.launches
    check_polyhedra_operations(engine)
    dev = set_algebra_answers(engine)
    assert engine.launches > before + 100
    assert dev == set_algebra_answers(OracleEngine())


def test_robust_avoid_three_levels_on_gpu(engine):
    from tests.test_multilevel_cpu import check_robust_avoid_end_to_end
    check_robust_avoid_end_to_end(engine, seeds=(3,))


def test_batched_state_machine_on_gpu(engine):
    """SURVEY.md 8f-2 on the device engine: solve(qpn, inits) for networks with children regroups the device calls."""
    from tests.test_multilevel_cpu import check_batched_state_machine
    check_batched_state_machine(engine)


def test_multilevel_worker_pool_on_gpu(engine):
    """Host worker processes served by ONE device engine (workers.py): identical to the one-process batch."""
    import qpn_b200
    rng = np.random.default_rng(12)
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    B = 21
    X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    one = qpn_b200.solve_multilevel_batch(net, X, engine)
    before = engine.launches
    with qpn_b200.MultilevelPool(net, 4, engine=engine) as pool:
        stats = {}
        par = pool.solve(X, stats=stats)
    assert engine.launches > before and stats["workers"] == 4 and stats["engine_calls"] > 0
    for a, b in zip(one, par):
        assert a["solved"] == b["solved"]
        assert np.array_equal(a["x_opt"], b["x_opt"]) if a["solved"] else np.array_equal(a["x_fail"], b["x_fail"])
    assert sum(r["solved"] for r in par) >= B - 2


def test_solve_qp_implicit_bounds_convexity_on_gpu(engine):
    """Row A8 on the device engine: solve_qp, the batched bound LPs of implicit_bounds, check_qp_convexity."""
    from tests.test_multilevel_cpu import check_qp_row_a8
    before = engine.launches
    check_qp_row_a8(engine)
    assert engine.launches > before + 8


def test_one_off_calls_with_plans_above_threshold(engine):
    """qpn_gavi_solve_batched / qpn_level_equilibrium_batched build one-off plans for batches >= 256."""
    import qpn_b200
    rng = np.random.default_rng(21)
    net, X = problems.ra_inits(rng, 320)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    dz = g["M"].shape[1]
    w = X[:, par]
    z0 = np.zeros((len(X), dz)); z0[:, :len(dec)] = X[:, dec]
    ret = engine.gavi_solve(g, w, z0)
    for k in range(0, len(X), 7):
        ro = cport.gavi_solve(g, z0[k], w[k])
        assert ro["status"] == ret["status"][k] and ro["pivots"] == ret["pivots"][k]
        assert np.array_equal(ro["z_full"], ret["z_full"][k]) and np.array_equal(ro["basis"], ret["basis"][k])
    views = [qpn_ref.node_view(net, p) for p in net.depth[3]]
    lv = qpn_b200.LevelArrays(net.n_vars, views, g, dec, par, max_iters=150, proj=None)
    ret = engine.level_equilibrium(lv, X)
    ro = cport.Level(net.n_vars, views, g, dec, par, 150, None).solve(X, threads=4)
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k


def test_all_rows_frozen_and_no_row_frozen(engine):
    """Corner cases of the frozen-row rule (oracle/avi_pivot.py: freeze), below and above the batch size at which
    one-off calls build plans: a system of free variables only (every row frozen, nothing is swept: an unconstrained
    strictly convex QP) and a box LCP (no free variable: nothing frozen)."""
    rng = np.random.default_rng(23)
    n = 6
    G = rng.normal(size=(n, n)); Q = G.T @ G + 0.5 * np.eye(n)
    g = dict(M=Q, N=np.eye(n), o=rng.normal(size=n), l1=np.full(n, -np.inf), u1=np.full(n, np.inf),
             A=np.zeros((0, n)), B=np.zeros((0, n)), l2=np.zeros(0), u2=np.zeros(0))
    for B in (5, 300):
        w = rng.normal(size=(B, n)); z0 = rng.normal(size=(B, n))
        ret = engine.gavi_solve(g, w, z0)
        assert (ret["status"] == 1).all()
        for k in range(0, B, max(1, B // 6)):
            ro = cport.gavi_solve(g, z0[k], w[k])
            assert ro["status"] == 1 and ro["pivots"] == ret["pivots"][k]
            assert np.array_equal(ro["z_full"], ret["z_full"][k]) and np.array_equal(ro["basis"], ret["basis"][k])
            assert np.allclose(Q @ ret["z"][k] + w[k] + g["o"], 0.0, atol=1e-9)
    M = Q + 0.3 * (G - G.T)
    for B in (4, 280):
        q = rng.normal(size=(B, n)); z0 = rng.uniform(-0.5, 1.5, (B, n))
        z, st, pv, bs = engine.avi_solve(M, q, np.zeros(n), np.ones(n), z0)
        zo, so, po, bo = cport.avi_solve_batched(M, q, np.zeros(n), np.ones(n), z0, threads=2)
        assert (so == 1).all() and np.array_equal(st, so) and np.array_equal(pv, po) and np.array_equal(bs, bo) and np.array_equal(z, zo)


# ---- edge cases and failure statuses through the C ABI --------------------------------------------
def _lifted(Q, c, A, l, u, z0):
    g = problems.qp_gavi(np.asarray(Q, float), np.asarray(c, float), np.asarray(A, float), np.asarray(l, float), np.asarray(u, float))
    avi = qpn_ref.convert(g)
    z0 = np.asarray(z0, float)
    return avi["M"], avi["o"], avi["l"], avi["u"], np.concatenate([z0, g["A"] @ z0])


def test_edge_cases_and_failure_statuses(engine):
    INF = np.inf
    cases = {
        "unbounded LP": _lifted(np.zeros((1, 1)), [-1.0], [[1.0]], [0.0], [INF], [1.0, 0.0]),
        "infeasible": _lifted(np.eye(1), [0.0], [[1.0], [1.0]], [-INF, 1.0], [-1.0, INF], [0.0, 0.0, 0.0]),
        "equality row": _lifted(np.eye(2), [1.0, -2.0], [[1.0, 1.0]], [1.0], [1.0], [0.0, 0.0, 0.0]),
        "duplicate rows": _lifted(np.zeros((2, 2)), [1.0, 1.0], [[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 1.0]], [0.0] * 4, [INF] * 4,
                                  [1.0, 2.0, 0, 0, 0, 0]),
        "degenerate vertex": _lifted(np.zeros((2, 2)), [1.0, 1.0], [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, -1.0]], [0.0, 0.0, 0.0, -INF],
                                     [INF, INF, INF, 0.0], [0.5, 0.7, 0, 0, 0, 0]),
    }
    seen = set()
    for name, (M, q, l, u, z0) in cases.items():
        zo, so, po, bo = cport.avi_solve(M, q, l, u, z0)
        z, s, p, b = engine.avi_solve(M, q[None], l, u, z0[None])
        assert s[0] == so and p[0] == po and np.array_equal(b[0], bo), name
        assert np.array_equal(z[0], zo, equal_nan=True), name
        seen.add(int(so))
    assert 1 in seen and len(seen) >= 2, seen                       # successes and at least one non-success code
    # plain box AVIs: n = 1, a fixed variable (l == u), bounds active at the start
    M = np.array([[2.0]]); q = np.array([[-3.0], [5.0], [0.5]]); l = np.array([0.0]); u = np.array([1.0])
    z0 = np.array([[0.0], [1.0], [0.25]])
    zo, so, po, bo = cport.avi_solve_batched(M, q, l, u, z0)
    z, s, p, b = engine.avi_solve(M, q, l, u, z0)
    assert np.array_equal(z, zo) and np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo)
    assert np.allclose(z[:, 0], [1.0, 0.0, 0.0])
    M = np.array([[1.0, 0.5], [0.5, 2.0]]); l = np.array([0.3, -INF]); u = np.array([0.3, INF])       # first variable fixed
    q = np.array([[1.0, -1.0]]); z0 = np.array([[0.3, 0.0]])
    zo, so, po, bo = cport.avi_solve(M, q[0], l, u, z0[0])
    z, s, p, b = engine.avi_solve(M, q, l, u, z0)
    assert so == 1 and bo[0] == 4 and np.array_equal(z[0], zo) and np.array_equal(b[0], bo) and p[0] == po
    # pivot budget exhausted -> MAX_ITERS from both
    net, g, avi, dec, par = problems.fp_avi()
    X, z0 = problems.fp_starts(np.random.default_rng(3), 4)
    qq = np.tile(avi["o"], (4, 1))
    zo, so, po, bo = cport.avi_solve_batched(avi["M"], qq, avi["l"], avi["u"], z0, max_pivots=5)
    z, s, p, b = engine.avi_solve(avi["M"], qq, avi["l"], avi["u"], z0, max_pivots=5)
    assert (so == 3).all() and np.array_equal(s, so) and np.array_equal(p, po)
    # empty batch is a no-op
    z, s, p, b = engine.avi_solve(avi["M"], np.zeros((0, 32)), avi["l"], avi["u"], np.zeros((0, 32)))
    assert z.shape == (0, 32) and len(s) == 0


def test_avi_larger_random_monotone(engine):
    """Sizes beyond the examples (n up to 96: three warps per instance), monotone AVIs M = G'G + skew."""
    rng = np.random.default_rng(33)
    for n in (40, 64, 96):
        B = 24
        Ms, qs, ls, us, z0s = [], [], [], [], []
        for _ in range(B):
            G = rng.normal(size=(n, n)) / np.sqrt(n); K = rng.normal(size=(n, n)) * 0.3
            Ms.append(G.T @ G + 0.05 * np.eye(n) + (K - K.T))
            qs.append(rng.normal(size=n))
            l = np.where(rng.uniform(size=n) < 0.3, -np.inf, -rng.uniform(0.1, 1.0, n))
            u = np.where(rng.uniform(size=n) < 0.3, np.inf, rng.uniform(0.1, 1.0, n))
            ls.append(l); us.append(u); z0s.append(rng.normal(size=n))
        Ms, qs, ls, us, z0s = map(np.array, (Ms, qs, ls, us, z0s))
        zo, so, po, bo = cport.avi_solve_batched(Ms, qs, ls, us, z0s, threads=8)
        z, s, p, b = engine.avi_solve(Ms, qs, ls, us, z0s)
        assert (so == 1).all(), (n, np.bincount(so))
        assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo), n
        assert np.array_equal(z, zo), n


def test_full_size_batches_by_properties(engine):
    """BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot run them in
    seconds): 65,536 four_player equilibria and 65,536 robust_avoid bottom-level equilibria."""
    import qpn_b200
    rng = np.random.default_rng(77)
    # four_player: strongly monotone game -> every start reaches the one equilibrium; solving again from it is a no-op
    net = qpn_b200.setup("four_player_matrix_game")
    solver = qpn_b200.BatchedSolver(net, engine=engine)
    B = 65536
    X = rng.uniform(-5, 5, (B, 8))
    ret = solver.solve_batch(X)
    assert ret["solved"].all() and (ret["iters"] == 2).all()
    assert np.ptp(ret["x"], axis=0).max() < 1e-9
    again = solver.solve_batch(ret["x"])
    assert again["solved"].all() and (again["iters"] == 1).all() and np.array_equal(again["x"], ret["x"])      # idempotent
    onet = examples.four_player_matrix_game()
    for k in rng.integers(0, B, 8):
        ro = qpn_ref.solve_level_bottom(onet, 1, X[k], solver.proj)
        assert np.array_equal(ro["x"], ret["x"][k]) and ro["pivots"] == ret["pivots"][k]
    solver.close()
    # robust_avoid bottom level (BASELINE configs[2] shape: perturbed obstacles): solved everywhere, idempotent, and a
    # sample agrees with the oracle bit for bit
    ra = qpn_b200.setup("robust_avoid_simple")
    rs = qpn_b200.BatchedSolver(ra, engine=engine)
    lv = rs.resident_level(3)
    Xr = np.tile(ra.default_initialization, (B, 1))
    Xr[:, 0:6] += 0.5 * rng.normal(size=(B, 6)); Xr[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    r1 = lv.solve(Xr)
    assert r1["solved"].all()
    assert np.array_equal(r1["x"][:, :12], Xr[:, :12])                       # the bottom level moves only s and eps
    r2 = lv.solve(r1["x"])
    assert r2["solved"].all() and (r2["iters"] == 1).all() and np.array_equal(r2["x"], r1["x"])
    pl = ra.network_depth_map[3]
    g, dec, par = qpn_b200.assembly.level_gavi(ra, pl)
    L = cport.Level(ra.n_vars, [qpn_b200.assembly.node_view(ra, p) for p in pl], g, dec, par, ra.options.max_iters, rs.proj)
    idx = rng.integers(0, B, 64)
    ro = L.solve(Xr[idx], threads=4)
    assert np.array_equal(ro["x"], r1["x"][idx]) and np.array_equal(ro["pivots"], r1["pivots"][idx])
    rs.close()


def test_verify_solution_fallback_accepted_on_device(engine):
    """qp_processing.jl:129-146: the least-squares multipliers have a negative entry, the sign-constrained fallback
    finds valid ones (how == 3) -- on the device, bit-equal to the oracle, for the row orders that reach the branch."""
    import itertools
    Qd, qd = np.zeros((2, 2)), np.array([1.0, 1.0])
    A = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, -1.0]])             # x >= 0, y >= 0, x - y >= 0, all active at 0
    l, u = np.zeros(3), np.full(3, np.inf)
    dec = np.array([0, 1], np.int32)
    seen = set()
    for perm in itertools.permutations(range(3)):
        view = (Qd, qd, A[list(perm)], l, u, dec)
        sol, lam, how, act = engine.verify_solution(view, np.zeros((3, 2)))
        so, lo, ho, ao = cport.verify_solution(*view, np.zeros(2))
        assert sol.all() and (how == ho).all() and np.array_equal(lam[0], lo) and np.array_equal(act[0], ao)
        assert np.allclose(A[list(perm)].T @ lam[0], qd, atol=1e-8) and (lam[0] > -1e-4).all()
        seen.add(int(ho))
    assert seen == {2, 3}, seen
    # the fallback's rejection (how == 4): the gradient is outside the cone of the active normals
    view = (Qd, np.array([-1.0, 0.5]), A, l, u, dec)
    sol, lam, how, act = engine.verify_solution(view, np.zeros((1, 2)))
    so, lo, ho, ao = cport.verify_solution(*view, np.zeros(2))
    assert (not sol[0]) and how[0] == ho == 4 and np.array_equal(lam[0], lo)


def test_verify_solution_square_singular_active_matrix(engine):
    """Known deviation: when the active matrix Abar is SQUARE and singular the reference's `Abar \\ qt`
    is a sparse LU that throws and falls through to the PATH branch (qp_processing.jl:114-115,129-146); here the
    rank-revealing QR returns a basic solution directly.  The `solution` flag agrees (a valid multiplier exists either
    way); lam / how may differ.  Pinned here: device == oracle, and the multiplier returned is a valid one."""
    Qd, qd = np.zeros((2, 3)), np.array([2.0, 2.0])
    A = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0]])               # x + y >= 0 (own constraint) and x + y <= 0 (a child's piece)
    l, u = np.array([0.0, -np.inf]), np.array([np.inf, 0.0])
    dec = np.array([0, 1], np.int32)
    view = (Qd, qd, A, l, u, dec)
    x = np.zeros((2, 3))
    sol, lam, how, act = engine.verify_solution(view, x)
    so, lo, ho, ao = cport.verify_solution(*view, x[0])
    assert sol.all() and so and how[0] == ho and np.array_equal(lam[0], lo) and np.array_equal(act[0], ao)
    assert list(act[0]) == [1, 2]                                    # pos_inds / neg_inds: Abar = [a, -a] is 2 x 2 and singular
    assert np.allclose(A[:, :2].T @ lam[0], qd, atol=1e-8) and lam[0][0] > -1e-4 and lam[0][1] < 1e-4
    # and a gradient outside the range of Abar is not a solution on either path
    view = (Qd, np.array([2.0, -2.0]), A, l, u, dec)
    sol, lam, how, act = engine.verify_solution(view, x)
    so, lo, ho, ao = cport.verify_solution(*view, x[0])
    assert (not sol.any()) and (not so) and how[0] == ho
