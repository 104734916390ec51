"""Base.in(x, poly) (sets.jl:820-853) at batch scale: npts points against npoly polyhedra in one launch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200
from oracle import cport
npts, npoly, m, d = 65536, 16, 20, 18
rng = np.random.default_rng(1)
polys = []
for _ in range(npoly):
    A = rng.normal(size=(m, d)) * (rng.uniform(size=(m, d)) < 0.5)
    polys.append((A, -rng.uniform(0.5, 3, m), rng.uniform(0.5, 3, m)))
x = rng.normal(size=(npts, d)) * 0.5
eng = qpn_b200.Engine(0)
for rep in range(3):
    t = time.time(); got = eng.halfspace_in(polys, x); dt = time.time() - t
print(f"{npts} points x {npoly} polys ({m} rows, d={d}): {dt*1e3:.2f} ms end to end with host buffers, inside fraction {got.mean():.3f}")
idx = rng.integers(0, npts, 200)
ok = all(got[j, p] == cport.halfspace_in(*polys[p], x[j], 1e-6) for j in idx for p in range(npoly))
print("sample of 3,200 memberships equals the oracle:", ok)
