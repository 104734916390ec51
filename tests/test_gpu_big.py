"""GPU parity of the global-memory tableau path (csrc/avi_pivot_big.cuh, csrc/qpn_big.cuh) -- the
engine for AVIs beyond the shared-memory tableau (lifted n up to 1,536).

Two kinds of cases: (1) the small parity suites of test_gpu_parity.py pushed through the big path
with the "force_big" option, so every branch (plans, per-instance matrices, CSC, presolve) is
compared bit for bit against the C oracle cheaply; (2) sizes that can ONLY run there."""
import numpy as np
import pytest

from oracle import cport, examples, qpn_ref
from tests import problems

pytestmark = pytest.mark.gpu


@pytest.fixture()
def big_engine(engine):
    engine.set_option("force_big", 1)
    yield engine
    engine.set_option("force_big", 0)


def monotone_gavi(rng, n, m, np_=0):
    """SURVEY.md 8d config 5 at a chosen size: Q = G'G + 0.1 I, rows of A normalised, l = A xbar - U(0.1,1), u = +inf."""
    G = rng.normal(size=(n, n)) / np.sqrt(n)
    Q = G.T @ G + 0.1 * np.eye(n)
    A = rng.normal(size=(m, n)); A /= np.linalg.norm(A, axis=1, keepdims=True)
    xbar = rng.normal(size=n)
    l = A @ xbar - rng.uniform(0.1, 1, m)
    g = problems.qp_gavi(Q, np.zeros(n), A, l, np.full(m, np.inf))
    return g, xbar


def test_big_avi_four_player_plan_dense_csc(big_engine):
    rng = np.random.default_rng(21)
    net, g, avi, dec, par = problems.fp_avi()
    B = 300
    X, z0 = problems.fp_starts(rng, B)
    q = np.tile(avi["o"], (B, 1))
    zo, so, po, bo = cport.avi_solve_batched(avi["M"], q, avi["l"], avi["u"], z0)
    z, s, p, b = big_engine.avi_solve(avi["M"], q, avi["l"], avi["u"], z0)           # shared matrix + bounds: plan
    assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo) and np.array_equal(z, zo)
    zc, sc, pc, bc = big_engine.avi_solve(None, q, avi["l"], avi["u"], z0, csc=problems.dense_to_csc(avi["M"], base=1), index_base=1)
    assert np.array_equal(zc, zo) and np.array_equal(sc, so) and np.array_equal(pc, po) and np.array_equal(bc, bo)
    # per-instance bounds: no plan, matrix read directly
    L, U = np.tile(avi["l"], (B, 1)), np.tile(avi["u"], (B, 1))
    z2, s2, p2, b2 = big_engine.avi_solve(avi["M"], q, L, U, z0)
    assert np.array_equal(z2, zo) and np.array_equal(s2, so) and np.array_equal(p2, po) and np.array_equal(b2, bo)
    z3, s3, p3, b3 = big_engine.avi_solve(None, q, L, U, z0, csc=problems.dense_to_csc(avi["M"], base=0), index_base=0)
    assert np.array_equal(z3, zo) and np.array_equal(s3, so) and np.array_equal(p3, po) and np.array_equal(b3, bo)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_big_avi_random_qps_per_instance_matrix(big_engine, kind):
    rng = np.random.default_rng(200 + kind)
    for trial in range(6):
        n = int(rng.integers(2, 6))
        Q, c, A, l, u, z0 = problems.random_qp(rng, kind, n=n)
        m = len(l)
        B = 16
        Ms, qs, ls, us, z0s = [], [], [], [], []
        for _ in range(B):
            Q, c, A, l, u, z0 = problems.random_qp(rng, kind, n=n, m=m - (n if kind != 1 else 0))
            g = problems.qp_gavi(Q, c, A, l, u)
            avi = qpn_ref.convert(g)
            Ms.append(avi["M"]); qs.append(avi["o"]); ls.append(avi["l"]); us.append(avi["u"]); z0s.append(np.concatenate([z0, g["A"] @ z0]))
        Ms, qs, ls, us, z0s = map(np.array, (Ms, qs, ls, us, z0s))
        zo, so, po, bo = cport.avi_solve_batched(Ms, qs, ls, us, z0s)
        z, s, p, b = big_engine.avi_solve(Ms, qs, ls, us, z0s)
        assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo) and np.array_equal(z, zo)


def test_big_gavi_example_levels(big_engine):
    rng = np.random.default_rng(22)
    cases = []
    net = examples.simple_bilevel()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[2], {})
    cases.append(("simple_bilevel L2", g, dec, par, rng.normal(size=(48, 4)) * 2))
    net, g, avi, dec, par = problems.fp_avi()
    cases.append(("four_player L1", g, dec, par, rng.uniform(-7, 7, (96, 8))))
    net, X = problems.ra_inits(rng, 96)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    cases.append(("robust_avoid L3", g, dec, par, X))
    for name, g, dec, par, X in cases:
        B = len(X)
        dz = g["M"].shape[1]
        w = X[:, par]
        z0 = np.zeros((B, dz)); z0[:, :len(dec)] = X[:, dec]
        ret = big_engine.gavi_solve(g, w, z0)                      # with plans (batch >= 2)
        one = big_engine.gavi_solve(g, w[:1], z0[:1])              # without
        for k in range(B):
            ro = cport.gavi_solve(g, z0[k], w[k])
            assert ro["status"] == ret["status"][k] and ro["pivots"] == ret["pivots"][k], name
            assert np.array_equal(ro["basis"], ret["basis"][k]) and np.array_equal(ro["z_full"], ret["z_full"][k]), name
        assert one["status"][0] == ret["status"][0] and one["pivots"][0] == ret["pivots"][0], name
        assert np.array_equal(one["z_full"][0], ret["z_full"][0]) and np.array_equal(one["basis"][0], ret["basis"][0]), name
        assert (ret["status"] == 1).all(), name


@pytest.mark.parametrize("n,m,B,feasible", [(40, 80, 24, False), (64, 128, 12, False), (64, 128, 24, True)])
def test_big_monotone_gavi_beyond_shared_memory(engine, n, m, B, feasible):
    """Lifted sizes 200 and 320: no shared-memory tableau exists for them (the second also exceeds one thread per row)."""
    rng = np.random.default_rng(23 + n)
    g, xbar = monotone_gavi(rng, n, m)
    O = rng.normal(size=(B, n))                                    # instances differ in the linear term (as parameters)
    g["N"] = np.eye(n); g["B"] = np.zeros((m, n))
    # infeasible starts: the presolve projection runs.  Feasible start: phase 0 of the plan leaves 64 free
    # variables without a pivot, which phase 1 must pick up exactly as the specification does.
    z0 = np.zeros((B, n + m)); z0[:, :n] = xbar + (0.0 if feasible else 1.0) * rng.normal(size=(B, n))
    before = engine.big_launches
    ret = engine.gavi_solve(g, O, z0)
    assert engine.big_launches > before, "this size must run on the global-memory tableau path"
    assert (ret["status"] == 1).all()
    for k in range(B):
        ro = cport.gavi_solve(g, z0[k], O[k])
        assert ro["status"] == ret["status"][k] and ro["pivots"] == ret["pivots"][k]
        assert np.array_equal(ro["basis"], ret["basis"][k])
        assert np.array_equal(ro["z_full"], ret["z_full"][k])
    # the solution is the unique minimiser of a strictly convex QP: KKT residuals
    x = ret["z"][:, :n]; lam = ret["z"][:, n:]
    Q, A = g["M"][:, :n], g["A"][:, :n]
    assert np.abs(x @ Q.T + O - lam @ A).max() < 1e-8
    assert (x @ A.T - g["l2"] > -1e-8).all() and (lam > -1e-10).all()
    assert np.abs(lam * (x @ A.T - g["l2"])).max() < 1e-7


def test_big_avi_direct_n300_shared_matrix(engine):
    """A plain AVI with n = 300 (a box-constrained strongly monotone LCP-like problem): plan path, CSR check."""
    rng = np.random.default_rng(29)
    n, B = 300, 10
    G = rng.normal(size=(n, n)) / np.sqrt(n)
    M = G.T @ G + 0.2 * np.eye(n) + 0.1 * (G - G.T)                 # monotone, not symmetric
    l = np.where(rng.uniform(size=n) < 0.3, -np.inf, -rng.uniform(0.1, 1, n))
    u = np.where(rng.uniform(size=n) < 0.3, np.inf, rng.uniform(0.1, 1, n))
    q = rng.normal(size=(B, n))
    z0 = rng.normal(size=(B, n))
    zo, so, po, bo = cport.avi_solve_batched(M, q, l, u, z0)
    z, s, p, b = engine.avi_solve(M, q, l, u, z0)
    assert (so == 1).all()
    assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo) and np.array_equal(z, zo)
    bad, r = engine.check_avi(M, q, l, u, z)
    assert (bad == 0).all()


def test_big_verify_solution(big_engine):
    from tests.test_gpu_parity import _verify_cases
    seen = set()
    for name, view, X in _verify_cases():
        sol, lam, how, act = big_engine.verify_solution(view, X)
        for k in range(len(X)):
            so, lo, ho, ao = cport.verify_solution(*view, X[k])
            assert so == sol[k] and ho == how[k], (name, k, ho, how[k])
            assert np.array_equal(ao, act[k]), (name, k)
            assert np.array_equal(lo, lam[k]), (name, k, lo, lam[k])
            seen.add(int(ho))
    assert {0, 2, 4} <= seen, f"verify_solution branches exercised: {seen}"


def test_big_level_examples(big_engine):
    import qpn_b200
    rng = np.random.default_rng(24)
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    proj = rng.normal(size=(4, 8))
    views = [qpn_ref.node_view(net, p) for p in net.depth[1]]
    la = qpn_b200.LevelArrays(8, views, g, dec, par, max_iters=150, proj=proj)
    X = np.vstack([rng.uniform(-5, 5, (40, 8)), rng.uniform(-8, 8, (40, 8))])
    ro = cport.Level(8, views, g, dec, par, 150, proj).solve(X)
    ret = big_engine.level_equilibrium(la, X)                                # one-off, plans (batch >= 2 on the big path)
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    lv = qpn_b200.ResidentLevel(big_engine, la)                              # resident, plans built by the big kernel
    ret = lv.solve(X)
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    lv.release()
    net, X = problems.ra_inits(rng, 48)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[3], {})
    views = [qpn_ref.node_view(net, p) for p in net.depth[3]]
    la = qpn_b200.LevelArrays(net.n_vars, views, g, dec, par, max_iters=150, proj=None)
    ro = cport.Level(net.n_vars, views, g, dec, par, 150, None).solve(X)
    ret = big_engine.level_equilibrium(la, X)
    assert ro["solved"].all()
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k


def single_node_net(rng, n, m):
    """A one-node QPNet at the shape of SURVEY.md 8d config 5: min 0.5 x'Qx + q'x  s.t.  l <= Ax."""
    g, xbar = monotone_gavi(rng, n, m)
    net = examples.Net(n)
    c = net.add_constraint(g["A"][:, :n], g["l2"], g["u2"])
    net.add_qp(g["M"][:, :n], rng.normal(size=n), [c], list(range(n)))
    net.add_edges([])
    return net, xbar


@pytest.mark.parametrize("n,m,B", [(40, 80, 12), (64, 128, 6)])
def test_big_single_node_level_beyond_shared_memory(engine, n, m, B):
    """verify_solution (QR of nd x k active rows in the global slot) + solve_qep (lifted 2n + 2m) + verify again, fused."""
    import qpn_b200
    rng = np.random.default_rng(31 + n)
    net, xbar = single_node_net(rng, n, m)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    views = [qpn_ref.node_view(net, 1)]
    proj = rng.normal(size=(3, n))
    X = xbar + rng.normal(size=(B, n)) * np.array([0.0 if k % 3 == 0 else 1.0 for k in range(B)])[:, None]
    la = qpn_b200.LevelArrays(n, views, g, dec, par, max_iters=50, proj=proj)
    before = engine.big_launches
    ret = engine.level_equilibrium(la, X)
    assert engine.big_launches > before
    ro = cport.Level(n, views, g, dec, par, 50, proj).solve(X, threads=4)
    assert ro["solved"].all() and (ro["iters"] == 2).all()
    for k in ("x", "iters", "pivots", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    assert np.array_equal(ret["solved"], ro["solved"])
    # the stand-alone entry point at the solution and at the start
    sol, lam, how, act = engine.verify_solution(views[0], np.vstack([ret["x"], X]))
    for k in range(2 * B):
        so, lo, ho, ao = cport.verify_solution(*views[0], np.vstack([ret["x"], X])[k])
        assert so == sol[k] and ho == how[k] and np.array_equal(ao, act[k]) and np.array_equal(lo, lam[k])
    assert sol[:B].all() and not sol[B:].all()


def test_monotone_stress_net_through_solve(engine):
    """setup(:monotone_stress) -> solve(qpn, inits): BASELINE.json configs[4] at a size the oracle checks in seconds."""
    import qpn_b200
    rng = np.random.default_rng(41)
    net = qpn_b200.setup("monotone_stress", n=48, m=96)
    B = 16
    X = net.default_initialization + rng.normal(size=(B, 48))
    res = qpn_b200.solve(net, X)
    g, dec, par = qpn_b200.assembly.level_gavi(net, [1])
    views = [qpn_b200.assembly.node_view(net, 1)]
    ro = cport.Level(48, views, g, dec, par, net.options.max_iters, qpn_b200.projection_vectors(net)).solve(X, threads=4)
    assert ro["solved"].all()
    for k in range(B):
        assert res[k]["solved"] and res[k]["pivots"] == ro["pivots"][k] and res[k]["iters"] == ro["iters"][k]
        assert np.array_equal(res[k]["x_opt"], ro["x"][k])
    # all starts reach the unique minimiser of the strictly convex QP
    assert np.ptp(np.array([r["x_opt"] for r in res]), axis=0).max() < 1e-8


def test_synthetic_chain_bottom_level(engine):
    """BASELINE.json configs[3]: the bottom level (node 3: 64 own variables, 136 parameters, lifted n = 256 of which 64
    rows are swept) as a resident level.  With plans that export the swept rows first it fits the shared-memory
    engine (thread per row); forced onto the global-memory engine it runs with the compact slot in shared memory and,
    with that switched off, with the slot in global memory.  All three: bit for bit the oracle."""
    import qpn_b200
    rng = np.random.default_rng(42)
    net = qpn_b200.setup("synthetic_chain")
    B = 24
    X = net.default_initialization + 0.7 * rng.normal(size=(B, net.n_vars))
    pl = net.network_depth_map[3]
    g, dec, par = qpn_b200.assembly.level_gavi(net, pl)
    views = [qpn_b200.assembly.node_view(net, p) for p in pl]
    proj = qpn_b200.projection_vectors(net)
    ro = cport.Level(net.n_vars, views, g, dec, par, net.options.max_iters, proj).solve(X, threads=4)
    assert ro["solved"].all()

    def check(ret):
        for k in ("x", "iters", "pivots", "lam"):
            assert np.array_equal(ret[k], ro[k]), k

    solver = qpn_b200.BatchedSolver(net, engine=engine)
    lv = solver.resident_level(3)
    info = lv.info()
    assert not info["big"] and info["n"] == 256 and info["plan_pivots"] > 0 and info["ncol0"] == 65
    check(lv.solve(X))
    solver.close()
    before = engine.big_launches
    engine.set_option("force_big", 1)
    try:
        solver = qpn_b200.BatchedSolver(net, engine=engine)
        lv = solver.resident_level(3)
        assert lv.info()["big"]
        check(lv.solve(X))                                   # compact slot in shared memory
        engine.set_option("big_slot_in_smem", 0)
        check(lv.solve(X))                                   # slot in global memory
        solver.close()
    finally:
        engine.set_option("big_slot_in_smem", 1)
        engine.set_option("force_big", 0)
    assert engine.big_launches >= before + 2


def test_big_edge_cases_and_failure_statuses(big_engine):
    """The edge cases of test_gpu_parity.py (unbounded, infeasible, equality rows, degenerate vertices, fixed
    variables, MAX_ITERS, empty batch) through the global-memory tableau path."""
    from tests.test_gpu_parity import test_edge_cases_and_failure_statuses as run
    before = big_engine.big_launches
    run(big_engine)
    assert big_engine.big_launches > before


def test_big_avi_larger_random_monotone_per_instance(big_engine):
    """Per-instance dense matrices (no plan): n = 40 and 96."""
    rng = np.random.default_rng(34)
    for n in (40, 96):
        B = 12
        Ms, qs, ls, us, z0s = [], [], [], [], []
        for _ in range(B):
            G = rng.normal(size=(n, n)) / np.sqrt(n); K = rng.normal(size=(n, n)) * 0.3
            Ms.append(G.T @ G + 0.05 * np.eye(n) + (K - K.T))
            qs.append(rng.normal(size=n))
            ls.append(np.where(rng.uniform(size=n) < 0.3, -np.inf, -rng.uniform(0.1, 1.0, n)))
            us.append(np.where(rng.uniform(size=n) < 0.3, np.inf, rng.uniform(0.1, 1.0, n)))
            z0s.append(rng.normal(size=n))
        Ms, qs, ls, us, z0s = map(np.array, (Ms, qs, ls, us, z0s))
        zo, so, po, bo = cport.avi_solve_batched(Ms, qs, ls, us, z0s, threads=4)
        z, s, p, b = big_engine.avi_solve(Ms, qs, ls, us, z0s)
        assert (so == 1).all()
        assert np.array_equal(s, so) and np.array_equal(p, po) and np.array_equal(b, bo) and np.array_equal(z, zo)
