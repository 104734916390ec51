"""Full three-level robust_avoid_simple solves for a batch of perturbed instances through solve(qpn, inits)
(host recursion per instance, every numeric step on the device, pieces memoised across the batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = qpn_b200.setup("robust_avoid_simple", seed=3)
rng = np.random.default_rng(0)
X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
t = time.time(); res = qpn_b200.solve(net, X); dt = time.time() - t
print(f"{B} full solves: {dt:.2f} s ({B/dt:.1f} equilibria/s), solved {np.mean([r['solved'] for r in res]):.3f}")
