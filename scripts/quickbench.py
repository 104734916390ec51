import sys, time; sys.path.insert(0,'.')
import numpy as np, qpn_b200
from oracle import examples, qpn_ref
from tests import problems
eng=qpn_b200.Engine(0)
rng=np.random.default_rng(0)
net,g,avi,dec,par=problems.fp_avi()
for B in (4096, 65536):
    X,z0=problems.fp_starts(rng,B); q=np.tile(avi["o"],(B,1))
    for rep in range(3):
        t=time.time(); z,s,p,b=eng.avi_solve(avi["M"],q,avi["l"],avi["u"],z0); dt=time.time()-t
    print("AVI FP host-call B",B,"ms",dt*1e3,"solves/s",B/dt,"ok",(s==1).all())
    lv=qpn_b200.LevelArrays(8,[qpn_ref.node_view(net,p) for p in net.depth[1]],g,dec,par,max_iters=150,proj=rng.normal(size=(4,8)))
    for rep in range(3):
        t=time.time(); ret=eng.level_equilibrium(lv,X); dt=time.time()-t
    print("LEVEL FP host-call B",B,"ms",dt*1e3,"eq/s",B/dt,"ok",ret["solved"].all(), "iters",np.median(ret["iters"]),"piv",np.median(ret["pivots"]))
net,X=problems.ra_inits(rng,8192)
g,dec,par=qpn_ref.level_gavi(net,net.depth[3],{})
lv=qpn_b200.LevelArrays(net.n_vars,[qpn_ref.node_view(net,p) for p in net.depth[3]],g,dec,par,max_iters=150,proj=None)
for rep in range(3):
    t=time.time(); ret=eng.level_equilibrium(lv,X); dt=time.time()-t
print("LEVEL RA-L3 B",len(X),"ms",dt*1e3,"eq/s",len(X)/dt,"ok",ret["solved"].mean(),"iters",np.median(ret["iters"]),"piv",np.median(ret["pivots"]))
