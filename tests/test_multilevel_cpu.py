"""CPU suite: the reference's ONLY test -- test/simple_bilevel.jl:4-21 (KAT-1..8) -- through the
multi-level host driver, the polyhedral operations and the solution-graph code.  The numeric
stand-in here is the CPU oracle (tests/oracle_engine.py); tests/test_gpu_parity.py runs the same
eight cases on the real engine."""
import json
import math
import os

import numpy as np
import pytest

import qpn_b200
from qpn_b200 import polyhedra as ph
from qpn_b200.model import INF, Poly
from tests.oracle_engine import OracleEngine

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check_simple_bilevel_kats(engine):
    kat = json.load(open(os.path.join(GOLDEN, "simple_bilevel_kat.json")))
    net = qpn_b200.setup(":simple_bilevel", gen_solution_map=True)            # test/simple_bilevel.jl:2
    ns = qpn_b200.NetSolver(net, engine)
    for w, X, s in zip(kat["W"], kat["X"], kat["S"]):
        ret = ns.solve(np.array(w + [0.0, 0.0]))                              # solve(qpn, [w; x0]), :18
        assert ret["solved"], ret.get("error")
        assert any(np.allclose(ret["x_opt"], w + xi, atol=1e-4) for xi in X), (w, ret["x_opt"])     # :19
        assert len(ret["Sol"][2]) >= s, (w, len(ret["Sol"][2]), s)                                 # :20


def test_simple_bilevel_known_answers():
    check_simple_bilevel_kats(OracleEngine())


def test_lower_level_solution_map_is_the_kink():
    """y = max(x, 0): the lower node's solution graph has the pieces {x <= 0, y = 0} and {x >= 0, y = x}."""
    net = qpn_b200.setup(":simple_bilevel", gen_solution_map=True)
    ns = qpn_b200.NetSolver(net, OracleEngine())
    ret = ns.solve_base(np.array([0.3, -0.2, 0.0, 0.0]), 2)                   # x = 0: both pieces meet
    assert ret["solved"]
    pieces = ret["Sol"][1]
    assert len(pieces) == 2
    for x, y, inside in [(-1.0, 0.0, True), (2.0, 2.0, True), (2.0, 0.0, False), (-1.0, 0.5, False), (1.0, 1.5, False)]:
        pt = np.array([0.0, 0.0, x, y])
        assert any(ph.contains(p, pt, tol=1e-9) for p in pieces) == inside, (x, y)


def check_polyhedra_operations(engine):
    """exemplar / isempty (sets.jl:591-655), issubset / remove_subsets (:377-407,889-902), complement (:918-975),
    project + simplify -- every LP through `engine`."""
    lp = ph.LPSolver(engine)
    box = Poly(np.eye(2), [0.0, 0.0], [1.0, 1.0])
    tri = Poly(np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]), [0.0, 0.0, -INF], [INF, INF, 1.0])
    assert ph.issubset(tri, box, lp) and not ph.issubset(box, tri, lp)
    assert len(ph.remove_subsets([box, tri], lp)) == 1
    assert not ph.isempty(box, lp) and ph.isempty(ph.intersect(box, Poly(np.array([[1.0, 0.0]]), [2.0], [INF])), lp)
    # open complement pieces: x in box  xor  x in some complement piece
    comp = ph.complement(box)
    assert len(comp) == 4
    for pt in ([0.5, 0.5], [1.5, 0.5], [0.0, 0.0], [1.0, 1.0 + 1e-3]):
        inside = ph.contains(box, np.array(pt), tol=0.0)
        assert inside != any(ph.contains(c, np.array(pt), tol=0.0) for c in comp), pt
    # an open set touching a closed one in a single point is empty (exemplar's dual-activity rule)
    touching = ph.intersect(box, Poly(np.array([[1.0, 0.0]]), [1.0], [INF], [True], [True], normalize=False))
    assert ph.isempty(touching, lp)
    # projection of the unit simplex in 3-D onto (x, y) is the triangle; of a line segment with an equality
    simplex = Poly(np.vstack([np.eye(3), np.ones((1, 3))]), [0, 0, 0, -INF], [INF, INF, INF, 1.0])
    pr = ph.project(simplex, [0, 1], lp)
    for pt, inside in ([0.2, 0.3], True), ([0.7, 0.7], False), ([-0.1, 0.2], False), ([0.5, 0.5], True):
        assert ph.contains(pr, np.array(pt), tol=1e-9) == inside, pt
    seg = Poly(np.array([[1.0, -1.0, 0.0], [0.0, 1.0, -1.0], [1.0, 0.0, 0.0]]), [0.0, 0.0, 0.0], [0.0, 0.0, 2.0])   # x=y=z in [0,2]
    pr = ph.project(seg, [2], lp)
    assert ph.contains(pr, np.array([1.5])) and not ph.contains(pr, np.array([2.5])) and not ph.contains(pr, np.array([-0.5]))
    # simplify merges the two one-sided rows that projection produces into one two-sided slice
    assert len(ph.simplify(pr)) == 1


def random_polytopes(seed, count, d=3):
    """Boxes cut by random halfspaces; every third one is a shrunken copy of its predecessor (a subset), every fifth one
    carries a contradictory pair of rows (empty)."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        if k % 3 == 2:
            prev = out[-1]
            pl = np.where(np.isinf(prev.l), prev.u - 1.0, prev.l)                # (a flipped row is bounded above only)
            pu = np.where(np.isinf(prev.u), prev.l + 1.0, prev.u)
            mid = 0.5 * (pl + pu)
            lo = np.where(np.isinf(prev.l), -INF, pl + 0.25 * (mid - pl))
            up = np.where(np.isinf(prev.u), INF, pu - 0.25 * (pu - mid))
            out.append(Poly(prev.A.copy(), lo, up, normalize=False))
            continue
        c = rng.normal(size=d)
        A = np.vstack([np.eye(d), rng.normal(size=(3, d))])
        lo = np.concatenate([c - rng.uniform(0.2, 1.5, d), A[d:] @ c - rng.uniform(0.1, 1.0, 3)])
        up = np.concatenate([c + rng.uniform(0.2, 1.5, d), np.full(3, INF)])
        if k % 5 == 4:
            a = rng.normal(size=(1, d))
            A = np.vstack([A, a, -a]); lo = np.concatenate([lo, [1.0, 1.0]]); up = np.concatenate([up, [INF, INF]])    # a x >= 1 and -a x >= 1
        out.append(Poly(A, lo, up))
    return out


def set_algebra_answers(engine, seed=11, count=14):
    """Every geometric predicate of the path on a family of random polytopes, as plain data (compared between engines)."""
    lp = ph.LPSolver(engine)
    P = random_polytopes(seed, count)
    empty = [bool(ph.isempty(p, lp)) for p in P]
    live = [p for p, e in zip(P, empty) if not e]
    sub = [[bool(ph.issubset(a, b, lp)) for b in live] for a in live]
    kept = len(ph.remove_subsets(live, lp))
    inter_empty = [[bool(ph.isempty(ph.intersect(a, b), lp)) for b in live] for a in live]
    proj_rows = [len(ph.project(p, [0, 1], lp)) for p in live[:6]]
    return dict(empty=empty, sub=sub, kept=kept, inter_empty=inter_empty, proj_rows=proj_rows)


def test_polyhedra_operations():
    check_polyhedra_operations(OracleEngine())
    ans = set_algebra_answers(OracleEngine())
    n = len(ans["sub"])
    assert all(ans["sub"][i][i] for i in range(n)) and ans["kept"] < n           # reflexive; the shrunken copies are dropped
    assert any(ans["empty"]) and not all(ans["empty"])
    assert sum(map(sum, ans["sub"])) > n                                         # some proper subsets


def test_solve_dispatch_keeps_result_fields():
    """requests.jl:18-22 / algorithm.jl:116,125: result field names of both outcomes."""
    net = qpn_b200.setup(":simple_bilevel")
    ns = qpn_b200.NetSolver(net, OracleEngine())
    ret = ns.solve(np.array([1.0, 0.0, 0.0, 0.0]))
    assert ret["solved"] and set(ret) >= {"solved", "x_opt", "Sol", "identified_request", "x_alts"}
    net.options.max_iters = 1                                               # cannot converge in one iteration from here
    ret = qpn_b200.NetSolver(net, OracleEngine()).solve(np.array([1.0, 0.0, 3.0, 0.0]))
    assert (not ret["solved"]) and ret["x_opt"] is None and "x_fail" in ret


def check_robust_avoid_end_to_end(engine, seeds=(3, 6)):
    """examples/robust_avoid_simple.jl: three levels (ego -> adversaries -> separating planes),
    exploration_vertices = 10.  With the stand-in problem data there is no reference
    output to compare with; the result must be a feasible point at which every node passes the
    reference's own optimality test against the solution pieces of its children."""
    for seed in seeds:
        net = qpn_b200.setup(":robust_avoid_simple", seed=seed)
        ns = qpn_b200.NetSolver(net, engine)
        ret = ns.solve(net.default_initialization)
        assert ret["solved"], (seed, ret.get("error"))
        x = ret["x_opt"]
        assert np.array_equal(x[:6], net.default_initialization[:6])          # xe, xo are parameters
        for P in net.constraints.values():
            assert ph.contains(P, x, tol=1e-6, closed=True)
        assert set(ret["Sol"]) == {1, 2, 3, 4, 5} and all(len(ret["Sol"][k]) >= 1 for k in (1, 2, 3, 4))
        # the bottom level alone: each separating-plane LP is at its optimum for the final (xe+ue, xo+uo)
        low = ns.solve_base(x, 3)
        assert low["solved"] and np.allclose(low["x_opt"], x, atol=1e-6)


def test_robust_avoid_three_levels():
    check_robust_avoid_end_to_end(OracleEngine())


def test_cycling_exit_is_reported_not_raised():
    """algorithm.jl:16-30,120-126: a repeated iterate ends the solve with solved=false."""
    net = qpn_b200.setup(":robust_avoid_simple", seed=5)
    ret = qpn_b200.NetSolver(net, OracleEngine()).solve(net.default_initialization)
    assert (not ret["solved"]) and "Cycling" in ret["error"] and ret["x_opt"] is None


def test_model_export_round_trip(tmp_path):
    """QPNet -> flat-array JSON -> QPNet gives the same arrays, graph and options, and the
    loaded net solves to the same answers (simple_bilevel KATs on the oracle stand-in)."""
    import json
    for name in ("simple_bilevel", "four_player_matrix_game", "robust_avoid_simple"):
        net = qpn_b200.setup(name)
        path = tmp_path / f"{name}.json"
        qpn_b200.export_net(net, path)
        doc = json.loads(path.read_text())                    # strict JSON: infinite bounds are null, never Infinity
        assert doc["format"] == "qpn-b200/1" and "Infinity" not in path.read_text()
        net2 = qpn_b200.load_net(path)
        assert net2.n_vars == net.n_vars and net2.network_depth_map == net.network_depth_map and net2.network_edges == net.network_edges
        assert net2.options == net.options and np.array_equal(net2.default_initialization, net.default_initialization)
        for i, qp in net.qps.items():
            q2 = net2.qps[i]
            assert np.array_equal(qp.Q, q2.Q) and np.array_equal(qp.q, q2.q) and qp.var_indices == q2.var_indices
            assert qp.constraint_indices == q2.constraint_indices
        for i, P in net.constraints.items():
            assert P == net2.constraints[i] and np.array_equal(P.A, net2.constraints[i].A) and np.array_equal(P.l, net2.constraints[i].l)
    net2 = qpn_b200.load_net(tmp_path / "simple_bilevel.json")
    ret = qpn_b200.NetSolver(net2, OracleEngine()).solve(np.array([1.0, 2.0, 0.0, 0.0]))
    ref = qpn_b200.NetSolver(qpn_b200.setup("simple_bilevel"), OracleEngine()).solve(np.array([1.0, 2.0, 0.0, 0.0]))
    assert ret["solved"] and np.array_equal(ret["x_opt"], ref["x_opt"])


def test_flatten_removes_edges_only():
    """programs.jl:117-124."""
    net = qpn_b200.setup("robust_avoid_simple")
    flat = qpn_b200.flatten(net)
    assert flat.network_depth_map == {1: [1, 2, 3, 4, 5]} and all(not v for v in flat.network_edges.values())
    assert net.num_levels() == 3                               # the original is untouched
    assert flat.decision_inds(5) == sorted(net.qps[5].var_indices)


def test_solve_qp_implicit_bounds_and_convexity():
    """SURVEY.md row A8 (qp_processing.jl:1-55, sets.jl:660-713) with the oracle as numeric stand-in."""
    check_qp_row_a8(OracleEngine())


def check_qp_row_a8(eng):
    from scipy.optimize import minimize
    rng = np.random.default_rng(5)
    for trial in range(6):
        n, m = 4, 7
        G = rng.normal(size=(n, n)); Q = G.T @ G + 0.1 * np.eye(n); q = rng.normal(size=n)
        A = rng.normal(size=(m, n)); xb = rng.normal(size=n)
        l = A @ xb - rng.uniform(0.1, 1, m); u = A @ xb + rng.uniform(0.1, 1, m)
        x = qpn_b200.solve_qp(eng, Q, q, A, l, u)
        assert (A @ x >= l - 1e-9).all() and (A @ x <= u + 1e-9).all()
        cons = [{"type": "ineq", "fun": lambda y, A=A, l=l: A @ y - l}, {"type": "ineq", "fun": lambda y, A=A, u=u: u - A @ y}]
        ref = minimize(lambda y: 0.5 * y @ Q @ y + q @ y, xb, constraints=cons, method="SLSQP", options=dict(ftol=1e-14, maxiter=500))
        assert abs((0.5 * x @ Q @ x + q @ x) - ref.fun) < 1e-7
    with pytest.raises(qpn_b200.qp.SolverFailure):
        qpn_b200.solve_qp(eng, np.zeros((1, 1)), [-1.0], [[1.0]], [0.0], [np.inf])          # unbounded LP
    # implicit equalities: x1 + x2 <= 1 and x1 + x2 >= 1 written as two one-sided rows; a true box row; a free direction
    A = np.array([[1.0, 1.0, 0.0], [-1.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    l = np.array([-np.inf, -np.inf, -2.0, 0.5]); u = np.array([1.0, -1.0, 2.0, 0.5])
    eq, vals = qpn_b200.implicit_bounds(eng, A, l, u)
    assert eq.tolist() == [True, True, False, True] and np.allclose(vals[[0, 1, 3]], [1.0, -1.0, 0.5])
    # Q indefinite, but positive on the null space of the implicit equality x1 + x2 = 1 (direction (1,-1,0)) and along x3
    Q = np.array([[1.0, 3.0, 0.0], [3.0, 1.0, 0.0], [0.0, 0.0, 2.0]])                      # eigenvalues 4, -2, 2; (1,-1,0) -> -2
    with pytest.raises(ValueError):
        qpn_b200.check_qp_convexity(eng, Q, A, l, u, [0, 1, 2], 7)
    Q2 = np.array([[1.0, -3.0, 0.0], [-3.0, 1.0, 0.0], [0.0, 0.0, 2.0]])                   # (1,1,0) -> -2 is excluded by the equality
    assert qpn_b200.check_qp_convexity(eng, Q2, A, l, u, [0, 1, 2], 7) > 0
    # the option reaches verify (qp_processing.jl:69): simple_bilevel is convex, the solve is unchanged
    net = qpn_b200.setup("simple_bilevel", check_convexity=True)
    ret = qpn_b200.NetSolver(net, eng).solve(np.array([1.0, 2.0, 0.0, 0.0]))
    ref = qpn_b200.NetSolver(qpn_b200.setup("simple_bilevel"), eng).solve(np.array([1.0, 2.0, 0.0, 0.0]))
    assert ret["solved"] and np.array_equal(ret["x_opt"], ref["x_opt"])


def test_projected_membership():
    """sets.jl:826-847: x in poly with only a prefix of the coordinates given (a feasibility QP over the rest)."""
    eng = OracleEngine()
    # {(x, y): 0 <= y <= 1, x - y = 0.5}: x is in the projection iff 0.5 <= x <= 1.5
    P = Poly(np.array([[0.0, 1.0], [1.0, -1.0]]), [0.0, 0.5], [1.0, 0.5])
    for x, want in ((0.4, False), (0.5, True), (1.0, True), (1.5, True), (1.6, False)):
        assert ph.contains_prefix(P, [x], eng) == want, x
    assert ph.contains_prefix(P, [1.0, 0.5], eng) and not ph.contains_prefix(P, [1.0, 0.6], eng)      # full dimension: plain membership
    # a lifted solution piece of simple_bilevel's lower node: (x, y, lam) with y = max(x, 0)
    Q = Poly(np.array([[1.0, -1.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), [0.0, 0.0, 0.0], [0.0, INF, 0.0])   # x = y >= 0, lam = 0
    assert ph.contains_prefix(Q, [2.0], eng) and not ph.contains_prefix(Q, [-1.0], eng)


def check_batched_state_machine(engine):
    """the batch form regroups the instances' device calls; results are those of the per-instance runs."""
    rng = np.random.default_rng(8)
    net = qpn_b200.setup("simple_bilevel")
    X = np.array([[w1, w2, 0.0, 0.0] for w1, w2 in rng.normal(size=(10, 2)) * 2])
    seq = [qpn_b200.NetSolver(net, engine).solve(x) for x in X]
    stats = {}
    bat = qpn_b200.solve_multilevel_batch(net, X, engine, stats=stats)
    assert all(a["solved"] and b["solved"] and np.array_equal(a["x_opt"], b["x_opt"]) for a, b in zip(seq, bat))
    assert stats["device_calls"] < stats["requests"]
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    B = 10
    X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    ns = qpn_b200.NetSolver(net, engine)
    seq = [ns.solve(x) for x in X]
    stats = {}
    bat = qpn_b200.solve_multilevel_batch(net, X, engine, chunk=6, stats=stats)            # two chunks: 6 + 4 threads
    for a, b in zip(seq, bat):
        assert a["solved"] == b["solved"] and (not a["solved"] or np.array_equal(a["x_opt"], b["x_opt"]))
    assert sum(b["solved"] for b in bat) >= B - 3 and stats["device_calls"] < stats["requests"]
    # the public entry point takes the batch form for networks with children
    out = qpn_b200.solve_multilevel_batch(net, X[:3], engine)
    assert [o["solved"] for o in out] == [b["solved"] for b in bat[:3]]


def test_batched_state_machine_matches_per_instance_runs():
    check_batched_state_machine(OracleEngine())


def test_batching_engine_propagates_errors_and_survives_early_exits():
    """A failing device call reaches every instance of its group as an exception; instances that finish early do
    not stall the others."""
    from qpn_b200.batching import BatchingEngine

    class Flaky(OracleEngine):
        def comp_indices(self, g, z, w, tol=1e-2):
            raise RuntimeError("device fault (injected)")

    be = BatchingEngine(Flaky())
    net = qpn_b200.setup("simple_bilevel")
    g, dec, par = qpn_b200.assembly.level_gavi(net, [1])

    def job(k):
        if k == 0:
            return "early"                                   # never touches the engine
        r = be.gavi_solve(g, np.zeros((1, g["N"].shape[1])), np.zeros((1, g["M"].shape[1])))
        if k == 1:
            return int(r["status"][0])
        try:
            be.comp_indices(g, r["z"], np.zeros((1, g["N"].shape[1])))
        except RuntimeError as e:
            return str(e)
        return "no error"

    out = be.run([lambda k=k: job(k) for k in range(5)])
    assert out[0] == "early" and out[1] == 1 and out[2:] == ["device fault (injected)"] * 3
    assert be.device_calls == 2 and be.requests == 7


def test_multilevel_batch_over_worker_processes():
    """A multi-level batch sharded over host processes served by ONE engine (workers.py): same results, in instance
    order, as the single-process batch; odd shard sizes; the pool and its memos serve several batches."""
    rng = np.random.default_rng(9)
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    B = 7
    X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    one = qpn_b200.solve_multilevel_batch(net, X, OracleEngine())
    with qpn_b200.MultilevelPool(net, 3, engine=OracleEngine()) as pool:
        for rep in range(2):
            stats = {}
            par = pool.solve(X, stats=stats)
            assert len(par) == B and stats["workers"] == 3 and 0 < stats["engine_calls"] <= stats["device_calls"] < stats["requests"]
            for a, b in zip(one, par):
                assert a["solved"] == b["solved"] and "Sol" not in b
                assert np.array_equal(a["x_opt"], b["x_opt"]) if a["solved"] else np.array_equal(a["x_fail"], b["x_fail"])
        kept = pool.solve(X[:2], keep_sol=True)                   # fewer instances than workers
        assert len(kept) == 2 and all("Sol" in r for r in kept if r["solved"])
    short = qpn_b200.solve_multilevel_workers(net, X[:3], 8, engine=OracleEngine())
    assert [r["solved"] for r in short] == [r["solved"] for r in one[:3]]


def test_worker_pool_reports_engine_errors():
    """An engine error reaches the caller of pool.solve as it reaches the caller of the one-process batch (an
    exception, not a hang), and the pool still shuts down."""
    class Broken(OracleEngine):
        def verify_solution(self, node, x, tol=1e-4):
            raise qpn_b200.EngineError("device lost")
    net = qpn_b200.setup("simple_bilevel")
    X = np.array([[1.0, 2.0, 0.0, 0.0], [0.5, -1.0, 0.0, 0.0], [2.0, 0.1, 0.0, 0.0]])
    with pytest.raises(Exception, match="device lost"):
        qpn_b200.solve_multilevel_batch(net, X, Broken())
    with qpn_b200.MultilevelPool(net, 2, engine=Broken()) as pool:
        with pytest.raises(RuntimeError, match="device lost"):
            pool.solve(X)
        with pytest.raises(RuntimeError, match="new pool"):
            pool.solve(X)


def test_multiplier_vertices_of_a_ten_row_node():
    """expand's get_verts (avi_solutions.jl:252-255) at a degenerate point: a node with 2 decision variables and 10
    constraint rows, 5 of them active at x = 0 (a pentagon's corner cut by three more lines through it).  The multiplier
    polytope { lam >= 0 on the active rows, A_d' lam = qt } has C(5, 2) candidate bases; its vertices are checked against
    brute-force enumeration, and the native enumeration (oracle build of csrc/net/vertex_enum.h, through the state
    machine) must lead to the same solution graph as the Python restatement."""
    import itertools
    from qpn_b200 import solgraph
    ang = np.array([0.3, 0.9, 1.4, 2.0, 2.6])
    A_act = np.stack([np.cos(ang), np.sin(ang)], 1)                 # five rows through the origin: a_i' x >= 0
    A_far = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0], [1.0, 1.0]])
    A = np.vstack([A_act, A_far]); l = np.concatenate([np.zeros(5), -np.full(5, 3.0)]); u = np.full(10, INF)
    qt = 0.7 * A_act[1] + 0.4 * A_act[3]                            # gradient in the cone of rows 1 and 3
    g = dict(M=np.hstack([np.zeros((2, 2)), -A.T]), N=np.zeros((2, 0)), o=qt, l1=np.full(2, -INF), u1=np.full(2, INF),
             A=np.hstack([A, np.zeros((10, 10))]), B=np.zeros((10, 0)), l2=l, u2=u)
    lam = np.zeros(10); lam[1], lam[3] = 0.7, 0.4
    z = np.concatenate([np.zeros(2), lam])
    verts = solgraph.multiplier_vertices(g, z, np.zeros(0), max_new=9)
    brute = []
    for i, j in itertools.combinations(range(5), 2):
        y = np.linalg.solve(A_act[[i, j]].T, qt)
        if (y >= -1e-9).all() and not (i, j) == (1, 3):
            brute.append((i, j, y))
    assert len(verts) == len(brute) >= 3
    for v, (i, j, y) in zip(verts, brute):                          # same (lexicographic) order, same multipliers
        assert np.allclose(v[2 + np.array([i, j])], y) and np.count_nonzero(np.abs(v[2:]) > 1e-12) <= 2
        assert np.allclose(A.T @ v[2:], qt)
    assert len(solgraph.multiplier_vertices(g, z, np.zeros(0), max_new=2)) == 2      # exploration_vertices caps the queue
    # a non-degenerate point (only two active rows) has nothing to explore
    z2 = z.copy()
    g2 = dict(g); g2["l2"] = np.concatenate([[-1.0, 0.0, -1.0, 0.0, -1.0], l[5:]])
    assert solgraph.multiplier_vertices(g2, z2, np.zeros(0), max_new=9) == []


def test_vertex_exploration_changes_the_solution_graph_and_native_agrees():
    """With exploration_vertices = 10 (examples/robust_avoid_simple.jl:4) degenerate nodes contribute the pieces of
    every vertex of their multiplier polytope: the bottom-level solution graphs of robust_avoid get larger than with
    exploration off, and the native enumeration reproduces the mirror's graphs row for row."""
    from tests.native_oracle import oracle_net, ra_inits, same_result
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    X = ra_inits(net, 24, seed=3)
    on = oracle_net(net).solve(X, keep_sol=True)
    eng, pieces, memo = OracleEngine(), {}, {}
    ref = [qpn_b200.NetSolver(net, eng, piece_cache=pieces, lp_memo=memo).solve(x) for x in X]
    assert all(same_result(a, b, sol=True) for a, b in zip(on, ref))
    net0 = qpn_b200.setup("robust_avoid_simple", seed=3, exploration_vertices=0)
    off = oracle_net(net0).solve(X, keep_sol=True)
    n_on = sum(len(r["Sol"][k]) for r in on if r["solved"] for k in (1, 2))
    n_off = sum(len(r["Sol"][k]) for r in off if r["solved"] for k in (1, 2))
    assert n_on > n_off > 0
