"""Wider parity sweep than the test suite has time for: resident levels of the bench against the C oracle, bit for bit,
on larger seeded samples.  usage: parity_sweep.py [scale]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
from oracle import cport
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = qpn_b200.Engine(0)
rng = np.random.default_rng(2024)
threads = len(os.sched_getaffinity(0))
bad = 0
for name, level, B in [("four_player_matrix_game", 1, 16384 * scale), ("robust_avoid_simple", 3, 4096 * scale), ("synthetic_chain", 3, 384 * scale),
                       ("synthetic_chain", 2, 0), ("monotone_stress", 1, 0)]:
    if B == 0:
        continue
    net = qpn_b200.setup(name)
    if name == "four_player_matrix_game":
        X = rng.uniform(-5, 5, (B, 8))
    elif name == "robust_avoid_simple":
        X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.5 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    else:
        X = net.default_initialization + 0.7 * rng.normal(size=(B, net.n_vars))
    solver = qpn_b200.BatchedSolver(net, engine=eng)
    lv = solver.resident_level(level)
    ret = lv.solve(X)
    pl = net.network_depth_map[level]
    g, dec, par = qpn_b200.assembly.level_gavi(net, pl)
    views = [qpn_b200.assembly.node_view(net, p) for p in pl]
    t = time.time()
    ro = cport.Level(net.n_vars, views, g, dec, par, net.options.max_iters, solver.proj).solve(X, threads=threads)
    dt = time.time() - t
    same = {k: bool(np.array_equal(ret[k], ro[k])) for k in ("x", "iters", "pivots", "lam", "solved")}
    nbad = int((~np.all(ret["x"] == ro["x"], axis=1)).sum())
    print(f"{name} level {level}: {B} instances, oracle {dt:.1f} s on {threads} threads, solved {ret['solved'].mean():.4f}, bit-equal {same}, instances with a differing x: {nbad}", flush=True)
    bad += 0 if all(same.values()) else 1
    solver.close()
print("levels with any difference:", bad)
sys.exit(1 if bad else 0)
