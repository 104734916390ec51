"""Full three-level robust_avoid_simple solves for a batch of perturbed instances: the per-instance recursion
run sequentially (pieces memoised) vs. through the BatchingEngine (device calls regrouped across instances)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = qpn_b200.setup("robust_avoid_simple", seed=3)
rng = np.random.default_rng(0)
X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
eng = qpn_b200.Engine(0)
ns = qpn_b200.NetSolver(net, eng)
l0 = eng.launches; t = time.time(); seq = [ns.solve(x) for x in X]; dt = time.time() - t
print(f"sequential: {B} full solves in {dt:.2f} s ({B/dt:.1f} equilibria/s), solved {np.mean([r['solved'] for r in seq]):.3f}, launches {eng.launches - l0}", flush=True)
st = {}
l0 = eng.launches; t = time.time(); res = qpn_b200.solve_multilevel_batch(net, X, eng, stats=st); dt = time.time() - t
same = all(a["solved"] == b["solved"] and (not a["solved"] or np.array_equal(a["x_opt"], b["x_opt"])) for a, b in zip(seq, res))
print(f"batched:    {B} full solves in {dt:.2f} s ({B/dt:.1f} equilibria/s), launches {eng.launches - l0}, {st}, identical to sequential: {same}")
