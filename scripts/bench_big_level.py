"""Config 5 of BASELINE.json as a one-node QPNet through the fused level loop on the global-memory
tableau path (verify -> solve_qep at lifted n = 2(n+m) -> verify).  usage: bench_big_level.py n m B [n_oracle]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200
from oracle import cport, qpn_ref
from tests.test_gpu_big import single_node_net
n, m, B = (int(a) for a in sys.argv[1:4])
n_oracle = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rng = np.random.default_rng(31 + n)
net, xbar = single_node_net(rng, n, m)
g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
views = [qpn_ref.node_view(net, 1)]
X = xbar + rng.normal(size=(B, n))
eng = qpn_b200.Engine(0)
la = qpn_b200.LevelArrays(n, views, g, dec, par, max_iters=50, proj=None)
t = time.time(); lv = qpn_b200.ResidentLevel(eng, la); print(f"upload + plans: {(time.time()-t)*1e3:.1f} ms", flush=True)
for rep in range(2):
    t = time.time(); ret = lv.solve(X); dt = time.time() - t
    print(f"n={n} m={m} lifted={2*(n+m)} B={B}: {dt*1e3:.1f} ms ({B/dt:.1f} equilibria/s) solved={ret['solved'].all()} iters={np.unique(ret['iters'])} pivots p50={int(np.median(ret['pivots']))}", flush=True)
for k in range(n_oracle):
    t = time.time()
    ro = cport.Level(n, views, g, dec, par, 50, None).solve(X[k:k+1])
    print(f"oracle[{k}]: {time.time()-t:.2f} s solved {ro['solved'][0]} pivots {ro['pivots'][0]} vs {ret['pivots'][k]}; x bit-equal {np.array_equal(ro['x'][0], ret['x'][k])} lam bit-equal {np.array_equal(ro['lam'][0], ret['lam'][k])}")
