import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200
from oracle import cport
from tests.test_gpu_big import monotone_gavi
n, m, B = (int(a) for a in sys.argv[1:4])
force = int(sys.argv[4])
rng = np.random.default_rng(5)
g, xbar = monotone_gavi(rng, n, m)
g["N"] = np.eye(n); g["B"] = np.zeros((m, n))
O = rng.normal(size=(B, n))
z0 = np.zeros((B, n + m)); z0[:, :n] = xbar
eng = qpn_b200.Engine(0)
eng.set_option("force_big", force)
ret = eng.gavi_solve(g, O, z0)
ok = True
for k in range(B):
    ro = cport.gavi_solve(g, z0[k], O[k])
    same = ro["status"] == ret["status"][k] and ro["pivots"] == ret["pivots"][k] and np.array_equal(ro["z_full"], ret["z_full"][k])
    ok &= same
print(n, m, B, "force", force, "status", ret["status"][:4], "pivots", ret["pivots"][:4], "parity", ok)
