"""Debug driver: the global-memory tableau path under the checking-barrier build (-DQPN_TRACE): reports CTAs whose
warps met at different barrier lines or arrived with partial warps."""
import os, sys, ctypes as C
os.environ["QPN_CUDA_LIB"] = os.path.abspath("quadraticprogramnetworks.jl_b200/lib/libqpn_cuda_trace.so")
sys.path.insert(0, '.')
import numpy as np, qpn_b200
from oracle import cport, qpn_ref
from tests.test_gpu_big import monotone_gavi, single_node_net
eng = qpn_b200.Engine(0)
lib = eng.lib
lib.qpn_trace_enable.restype = C.POINTER(C.c_int)
lib.qpn_trace_enable.argtypes = [C.c_void_p, C.c_int]
def report(tr, B, label):
    a = np.ctypeslib.as_array(tr, shape=(B * 16,)).reshape(B, 16).copy()
    hit = np.nonzero((a[:, :4] != -1).any(axis=1))[0]
    print(label, "CTAs with records:", len(hit), [tuple(a[b, :4]) for b in hit[:6]], flush=True)
    return len(hit)
rng = np.random.default_rng(5)
bad = 0
for (n, m, B, feas) in [(40, 80, 64, False), (64, 128, 148, True), (64, 128, 148, False)]:
    g, xbar = monotone_gavi(rng, n, m)
    g["N"] = np.eye(n); g["B"] = np.zeros((m, n))
    O = rng.normal(size=(B, n))
    z0 = np.zeros((B, n + m)); z0[:, :n] = xbar + (0.0 if feas else 1.0) * rng.normal(size=(B, n))
    tr = lib.qpn_trace_enable(eng.h, max(B, 1024))
    ret = eng.gavi_solve(g, O, z0)
    ro = cport.gavi_solve(g, z0[3], O[3])
    print("gavi", n, m, B, "ok", (ret["status"] == 1).all(), "parity", np.array_equal(ro["z_full"], ret["z_full"][3]))
    bad += report(tr, max(B, 1024), "gavi big")
for (n, m, B) in [(40, 80, 48), (64, 128, 148)]:
    net, xbar = single_node_net(rng, n, m)
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    views = [qpn_ref.node_view(net, 1)]
    X = xbar + rng.normal(size=(B, n))
    la = qpn_b200.LevelArrays(n, views, g, dec, par, max_iters=50, proj=rng.normal(size=(3, n)))
    tr = lib.qpn_trace_enable(eng.h, max(B, 1024))
    ret = eng.level_equilibrium(la, X)
    print("level", n, m, B, "solved", ret["solved"].all())
    bad += report(tr, max(B, 1024), "level big")
eng.set_option("force_big", 1)
from tests import problems
net, g, avi, dec, par = problems.fp_avi()
X, z0 = problems.fp_starts(rng, 300)
tr = lib.qpn_trace_enable(eng.h, 1024)
z, s, p, b = eng.avi_solve(avi["M"], np.tile(avi["o"], (300, 1)), avi["l"], avi["u"], z0)
bad += report(tr, 1024, "avi big (four_player)")
print("TOTAL records:", bad)
