"""GPU suite at the sizes BASELINE.json names: the 65,536-instance robust_avoid bottom-level
batch and one wave of the n = 256 / m = 512 stress shape against the C oracle, and the size-independent properties of
the full three-level batch (determinism, batch-composition independence, feasibility of every reported equilibrium)."""
import os

import numpy as np
import pytest

import qpn_b200
from oracle import cport
from tests.native_oracle import oracle_net

pytestmark = pytest.mark.gpu
THREADS = len(os.sched_getaffinity(0))


def test_robust_avoid_bottom_level_65536_bit_equal(engine):
    """BASELINE configs[2]'s batch size on the fused level kernel (level 3 of 3): x, lam, iteration and pivot counts,
    statuses -- all bit-equal to the oracle's level loop."""
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    B = 65536
    X = qpn_b200.examples.robust_avoid_batch(net, B, seed=42)
    solver = qpn_b200.BatchedSolver(net, engine=engine)
    lv = solver.resident_level(3)
    ret = lv.solve(X)
    pl = net.network_depth_map[3]
    g, dec, par = qpn_b200.assembly.level_gavi(net, pl)
    views = [qpn_b200.assembly.node_view(net, p) for p in pl]
    ro = cport.Level(net.n_vars, views, g, dec, par, net.options.max_iters, solver.proj).solve(X, threads=THREADS)
    for k in ("solved", "iters", "pivots", "x", "lam"):
        assert np.array_equal(ret[k], ro[k]), k
    assert ret["solved"].mean() > 0.99
    solver.close()


def test_monotone_stress_n256_m512_one_wave(engine):
    """BASELINE configs[4] at full size (lifted level AVI n = 1,536, global-memory tableau path): one wave of 148
    instances on the device; a sample of them against the oracle bit for bit (the oracle needs ~4 s per instance)."""
    net = qpn_b200.setup("monotone_stress")
    B = 148
    X = net.default_initialization + np.random.default_rng(9).normal(size=(B, net.n_vars))
    solver = qpn_b200.BatchedSolver(net, engine=engine)
    lv = solver.resident_level(1)
    assert lv.info()["big"] and lv.info()["n"] == 1536
    ret = lv.solve(X)
    assert ret["solved"].all()
    pick = [0, 37, 73, 111, 147][: max(2, min(5, THREADS // 3))]
    g, dec, par = qpn_b200.assembly.level_gavi(net, [1])
    views = [qpn_b200.assembly.node_view(net, 1)]
    ro = cport.Level(net.n_vars, views, g, dec, par, net.options.max_iters, solver.proj).solve(X[pick], threads=len(pick))
    for k in ("solved", "iters", "pivots", "x", "lam"):
        assert np.array_equal(ret[k][pick], ro[k]), k
    # KKT residuals of every instance of the wave (a property the size does not change)
    qp, P = net.qps[1], net.constraints[1]
    x, lam = ret["x"], ret["lam"]
    # (rows whose leading coefficient was negative are stored flipped -- sets.jl:76-89 -- so a row may be bounded above
    # and its multiplier negative)
    ax = x @ P.A.T
    # (the solve accepts its own AVI residual at 1e-6 per component -- avi.jl:148-156 -- and verify_solution re-derives lam
    # by least squares, so 1e-5 is the scale to expect; the measured maximum on this wave is 1.1e-6)
    assert np.abs(x @ qp.Q.T + qp.q - lam @ P.A).max() < 1e-5
    assert (ax >= P.l - 1e-6).all() and (ax <= P.u + 1e-6).all()
    lower_only, upper_only = np.isinf(P.u) & ~np.isinf(P.l), np.isinf(P.l) & ~np.isinf(P.u)
    assert lam[:, lower_only].min(initial=0.0) > -1e-6 and lam[:, upper_only].max(initial=0.0) < 1e-6
    gap = np.where(lower_only, ax - P.l, np.where(upper_only, P.u - ax, 0.0))
    assert np.abs(lam * gap).max() < 1e-5                            # complementarity
    solver.close()


def test_robust_avoid_three_levels_full_batch_properties(engine):
    """The 65,536-instance three-level batch (BASELINE configs[2]) on the device: identical on a second run,
    independent of the batch an instance travels in (a slice re-solved alone, a sample re-solved on the oracle build),
    and every reported equilibrium is feasible for every constraint of the network."""
    from qpn_b200.netsolve import NetBinding
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    B = 65536
    X = qpn_b200.examples.robust_avoid_batch(net, B, seed=0)
    nb = NetBinding(net, engine.lib, "qpn_net_", handle=engine.h, threads=6)
    a = nb.solve_arrays(X)
    b = nb.solve_arrays(X)
    for k in ("solved", "level_iters", "error", "x"):
        assert np.array_equal(a[k], b[k]), k
    assert a["solved"].mean() > 0.9
    part = nb.solve_arrays(X[30000:31000])
    for k in ("solved", "level_iters", "error", "x"):
        assert np.array_equal(a[k][30000:31000], part[k]), k
    sample = np.random.default_rng(1).choice(B, 2048, replace=False)
    ref = oracle_net(net, threads=THREADS).solve_arrays(X[sample])
    for k in ("solved", "level_iters", "error", "x"):
        assert np.array_equal(a[k][sample], ref[k]), k
    xs = a["x"][a["solved"]]
    assert np.array_equal(xs[:, :6], X[a["solved"]][:, :6])               # xe, xo are parameters: nobody moves them
    for P in net.constraints.values():
        ax = xs @ P.A.T
        assert (ax >= P.l - 1e-6).all() and (ax <= P.u + 1e-6).all()
    nb.close()
