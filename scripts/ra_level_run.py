"""robust_avoid bottom level (players s1, s2; n = 52 lifted AVI) on a resident level: B instances."""
import sys; sys.path.insert(0,'.')
import numpy as np, qpn_b200
from qpn_b200 import assembly
B=int(sys.argv[1]) if len(sys.argv)>1 else 8192
eng=qpn_b200.Engine(0)
rng=np.random.default_rng(0)
ra=qpn_b200.setup("robust_avoid_simple")
gr,decr,parr=assembly.level_gavi(ra,[1,2])
lar=qpn_b200.LevelArrays(18,[assembly.node_view(ra,p) for p in (1,2)],gr,decr,parr,150,None)
X=np.tile(ra.default_initialization,(B,1)); X[:,0:6]+=0.5*rng.normal(size=(B,6)); X[:,6:12]=rng.uniform(-1,1,(B,6))
lv=qpn_b200.ResidentLevel(eng,lar)
for _ in range(3): r=lv.solve(X,want_lam=False)
print("solved",r["solved"].mean(),"pivots p50",np.median(r["pivots"]),"iters",np.bincount(r["iters"]))
