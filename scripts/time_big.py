"""Timing breakdown of the global-memory tableau path: plan vs. solve, per batch size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200
from tests.test_gpu_big import monotone_gavi
n, m = int(sys.argv[1]), int(sys.argv[2])
Bs = [int(a) for a in sys.argv[3:]]
rng = np.random.default_rng(5)
g, xbar = monotone_gavi(rng, n, m)
g["N"] = np.eye(n); g["B"] = np.zeros((m, n))
eng = qpn_b200.Engine(0)
ga = qpn_b200.engine.GaviArrays(g)
for B in Bs:
    O = rng.normal(size=(B, n))
    z0 = np.zeros((B, n + m)); z0[:, :n] = xbar
    for rep in range(2):
        t = time.time(); ret = eng.gavi_solve(ga, O, z0); dt = time.time() - t
    piv = ret["pivots"].astype(float).sum()
    print(f"n={n} m={m} B={B}: {dt*1e3:.1f} ms, {B/dt:.1f} solves/s, pivots p50 {int(np.median(ret['pivots']))}, ok {(ret['status']==1).all()}", flush=True)
