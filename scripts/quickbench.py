"""Ad-hoc timings of the C-ABI entry points (host clock around synchronous calls, pageable buffers)."""
import sys, time; sys.path.insert(0,'.')
import numpy as np, qpn_b200
from qpn_b200 import assembly
eng=qpn_b200.Engine(0)
rng=np.random.default_rng(0)
def timeit(f, reps=5):
    f(); ts=[]
    for _ in range(reps):
        t=time.perf_counter(); r=f(); ts.append(time.perf_counter()-t)
    return min(ts), r
fp=qpn_b200.setup("four_player_matrix_game")
g,dec,par=assembly.level_gavi(fp,[1,2,3,4])
la=qpn_b200.LevelArrays(8,[assembly.node_view(fp,p) for p in (1,2,3,4)],g,dec,par,150,qpn_b200.projection_vectors(fp))
ra=qpn_b200.setup("robust_avoid_simple")
gr,decr,parr=assembly.level_gavi(ra,[1,2])
lar=qpn_b200.LevelArrays(18,[assembly.node_view(ra,p) for p in (1,2)],gr,decr,parr,150,None)
def ra_inits(B):
    X=np.tile(ra.default_initialization,(B,1)); X[:,0:6]+=0.5*rng.normal(size=(B,6)); X[:,6:12]=rng.uniform(-1,1,(B,6)); return X
for B in (4096,65536):
    X=rng.uniform(-5,5,(B,8))
    z0=np.zeros((B,24)); z0[:,:8]=X
    dt,r=timeit(lambda: eng.gavi_solve(g,np.zeros((B,0)),z0)); print(f"gavi_solve FP one-off   B={B}: {dt*1e3:.3f} ms  {B/dt/1e6:.2f} M/s ok={(r['status']==1).all()}")
    dt,r=timeit(lambda: eng.level_equilibrium(la,X,want_lam=False)); print(f"level FP one-off        B={B}: {dt*1e3:.3f} ms  {B/dt/1e6:.2f} M/s ok={r['solved'].all()}")
    lv=qpn_b200.ResidentLevel(eng,la)
    dt,r=timeit(lambda: lv.solve(X,want_lam=False)); print(f"level FP resident       B={B}: {dt*1e3:.3f} ms  {B/dt/1e6:.2f} M/s ok={r['solved'].all()}")
    lv.release()
for B in (8192,65536):
    X=ra_inits(B)
    dt,r=timeit(lambda: eng.level_equilibrium(lar,X,want_lam=False)); print(f"level RA-L3 one-off     B={B}: {dt*1e3:.3f} ms  {B/dt/1e6:.2f} M/s ok={r['solved'].mean()}")
    lv=qpn_b200.ResidentLevel(eng,lar)
    dt,r=timeit(lambda: lv.solve(X,want_lam=False)); print(f"level RA-L3 resident    B={B}: {dt*1e3:.3f} ms  {B/dt/1e6:.2f} M/s ok={r['solved'].mean()} piv p50={np.median(r['pivots'])}")
    lv.release()
