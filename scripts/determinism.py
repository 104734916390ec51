"""Race surrogate (compute-sanitizer is not available on the pool): every resident level of the bench, many launches on
the same inputs, outputs compared bit for bit with the first launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
eng = qpn_b200.Engine(0)
rng = np.random.default_rng(3)
bad = 0
for name, level, B, init in [("four_player_matrix_game", 1, 8192, lambda net, B: rng.uniform(-5, 5, (B, 8))),
                             ("robust_avoid_simple", 3, 8192, None), ("synthetic_chain", 3, 1184, None), ("monotone_stress", 1, 148, None)]:
    net = qpn_b200.setup(name)
    if init is not None:
        X = init(net, B)
    elif name == "robust_avoid_simple":
        X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.5 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    else:
        X = net.default_initialization + 0.7 * rng.normal(size=(B, net.n_vars))
    solver = qpn_b200.BatchedSolver(net, engine=eng)
    lv = solver.resident_level(level)
    first = lv.solve(X)
    n_rep = reps if name != "monotone_stress" else max(2, reps // 10)
    diff = 0
    for _ in range(n_rep):
        r = lv.solve(X)
        diff += int(not all(np.array_equal(r[k], first[k]) for k in ("x", "iters", "pivots", "solved", "lam")))
    print(f"{name}: {n_rep} repeats of {B} instances, launches that differ from the first: {diff}, solved {first['solved'].mean():.4f}", flush=True)
    bad += diff
    solver.close()
print("TOTAL differing launches:", bad)
sys.exit(1 if bad else 0)
