"""cProfile of the one-process multi-level batch on the device engine (where does the host time go?)."""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = qpn_b200.setup("robust_avoid_simple", seed=3)
rng = np.random.default_rng(0)
X = np.tile(net.default_initialization, (2 * B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(2 * B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (2 * B, 6))
eng = qpn_b200.Engine(0)
ns = qpn_b200.NetSolver(net, eng)
for x in X[:32]:
    ns.solve(x)
t = time.time(); [ns.solve(x) for x in X[:B]]; print(f"sequential {B / (time.time() - t):.1f}/s")
pr = cProfile.Profile(); pr.enable(); [ns.solve(x) for x in X[B:]]; pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
