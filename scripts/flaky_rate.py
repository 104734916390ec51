"""Debug: failure rate of [FP avi 65536 -> RA gavi 8192] for a given library build."""
import os, sys, subprocess
code = r'''
import os, sys
sys.path.insert(0, '.')
import numpy as np, qpn_b200
from oracle import examples, qpn_ref
from tests import problems
eng = qpn_b200.Engine(0)
rng = np.random.default_rng(0)
net, g, avi, dec, par = problems.fp_avi()
fp = {}
for BB in (4096, 65536):
    X, z0 = problems.fp_starts(rng, BB); proj = rng.normal(size=(4, 8)); fp[BB] = (X, z0, proj)
netr, XR = problems.ra_inits(rng, 8192)
gr, decr, parr = qpn_ref.level_gavi(netr, netr.depth[3], {})
B = 8192
X, z0, proj = fp[65536]; q = np.tile(avi["o"], (65536, 1))
dz = gr["M"].shape[1]; w = XR[:, parr]; z0r = np.zeros((B, dz)); z0r[:, :len(decr)] = XR[:, decr]
n_ok = 0
for attempt in range(int(sys.argv[1])):
    if sys.argv[2] == "fpfirst": eng.avi_solve(avi["M"], q, avi["l"], avi["u"], z0)
    r = eng.gavi_solve(gr, w, z0r); assert (r["status"] == 1).all()
    n_ok += 1
    print("ok", n_ok, flush=True)
'''
open("/tmp/_fl.py", "w").write(code)
for lib, mode, reps in [("libqpn_cuda.so", "fpfirst", 6), ("libqpn_cuda.so", "raonly", 12)]:
    fails = 0; total_ok = 0
    for proc in range(4):
        env = dict(os.environ, QPN_CUDA_LIB=os.path.abspath("quadraticprogramnetworks.jl_b200/lib/" + lib))
        r = subprocess.run([sys.executable, "/tmp/_fl.py", str(reps), mode], capture_output=True, text=True, env=env, timeout=300)
        oks = r.stdout.count("ok ")
        total_ok += oks
        if r.returncode != 0: fails += 1
    print(lib, mode, "processes failed:", fails, "of 4; successful calls:", total_ok, flush=True)
