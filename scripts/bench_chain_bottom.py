"""BASELINE.json configs[3] (synthetic 3-level chain, n = 64 per node): its bottom level (node 3: 64 own variables,
136 parameters, lifted level AVI n = 256) as a resident level.  usage: bench_chain_bottom.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
net = qpn_b200.setup("synthetic_chain")
solver = qpn_b200.BatchedSolver(net)
t = time.time(); lv = solver.resident_level(3); info = lv.info(); print(f"upload + plans {1e3*(time.time()-t):.1f} ms, info {info}", flush=True)
rng = np.random.default_rng(42)
if os.environ.get("QPN_SMEM_THREADS"):
    solver.engine.set_option("big_smem_threads", int(os.environ["QPN_SMEM_THREADS"]))
if os.environ.get("QPN_SLOT_IN_SMEM"):
    solver.engine.set_option("big_slot_in_smem", int(os.environ["QPN_SLOT_IN_SMEM"]))
for B in [int(a) for a in sys.argv[1:]] or [1024]:
    X = net.default_initialization + 0.7 * rng.normal(size=(B, net.n_vars))
    lv.solve(X[:148])
    t = time.time(); ret = lv.solve(X); dt = time.time() - t
    print(f"B={B}: {1e3*dt:.1f} ms ({B/dt:.0f} equilibria/s) solved {ret['solved'].mean():.4f} iters {np.unique(ret['iters'])} pivots p50 {int(np.median(ret['pivots']))}", flush=True)
