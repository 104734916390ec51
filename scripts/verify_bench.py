"""verify_solution (KKT dual recovery, qp_processing.jl:57-149) on its own: timing + a workload for ncu.
usage: verify_bench.py small|big"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200
mode = sys.argv[1] if len(sys.argv) > 1 else "small"
eng = qpn_b200.Engine(0)
rng = np.random.default_rng(3)
if mode == "small":
    # robust_avoid bottom-level node (nd = 3, m = 10) at its equilibria: active rows -> QR every instance
    ra = qpn_b200.setup("robust_avoid_simple")
    solver = qpn_b200.BatchedSolver(ra, engine=eng)
    B = 65536
    X = np.tile(ra.default_initialization, (B, 1)); X[:, 0:6] += 0.5 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    X = solver.resident_level(3).solve(X)["x"]
    view = qpn_b200.assembly.node_view(ra, 1)
else:
    # config 5 node (nd = 256, m = 512) at its equilibria: QR of 256 x k_active in the global slot
    ms = qpn_b200.setup("monotone_stress")
    solver = qpn_b200.BatchedSolver(ms, engine=eng)
    B = 148
    X = solver.resident_level(1).solve(ms.default_initialization + rng.normal(size=(B, ms.n_vars)))["x"]
    view = qpn_b200.assembly.node_view(ms, 1)
na = qpn_b200.NodeArrays(*view)
for rep in range(3):
    t = time.time(); sol, lam, how, act = eng.verify_solution(na, X); dt = time.time() - t
print(f"{mode}: nd={na.nd} m={na.m} B={B}: {dt*1e3:.2f} ms end to end ({B/dt:.0f} verifies/s), solutions {sol.mean():.3f}, active rows p50 {np.median((act != 0).sum(axis=1))}, how {np.bincount(how)}")
