"""Where does a full three-level robust_avoid_simple solve spend its time on the host mirror?"""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
net = qpn_b200.setup("robust_avoid_simple", seed=seed)
eng = qpn_b200.Engine(0)
ns = qpn_b200.NetSolver(net, eng)
rng = np.random.default_rng(0)
for k in range(3):
    x0 = net.default_initialization.copy()
    if k:
        x0[0:6] += 0.5 * rng.normal(size=6); x0[6:12] = rng.uniform(-1, 1, 6)
    l0 = eng.launches; t = time.time(); r = ns.solve(x0); dt = time.time() - t
    print(f"instance {k}: {dt:.2f} s, solved {r['solved']}, launches {eng.launches - l0}, err {r.get('error')}", flush=True)
pr = cProfile.Profile(); pr.enable(); ns.solve(net.default_initialization); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
