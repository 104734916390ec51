"""Full three-level robust_avoid_simple solves for a batch of perturbed instances (BASELINE.json configs[2]):
the per-instance recursion run sequentially (pieces memoised), through the BatchingEngine (device calls regrouped
across instances), and sharded over host worker processes that share the GPU.
usage: ra_full_batch.py [B] [workers ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    worker_counts = [int(a) for a in sys.argv[2:]]
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    rng = np.random.default_rng(0)
    X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    eng = qpn_b200.Engine(0)
    ns = qpn_b200.NetSolver(net, eng)
    Bs = min(B, 128)
    l0 = eng.launches; t = time.time(); seq = [ns.solve(x) for x in X[:Bs]]; dt = time.time() - t
    print(f"sequential: {Bs} full solves in {dt:.2f} s ({Bs/dt:.1f} equilibria/s), solved {np.mean([r['solved'] for r in seq]):.3f}, launches {eng.launches - l0}", flush=True)
    st = {}
    l0 = eng.launches; t = time.time(); res = qpn_b200.solve_multilevel_batch(net, X, eng, stats=st); dt = time.time() - t
    same = all(a["solved"] == b["solved"] and (not a["solved"] or np.array_equal(a["x_opt"], b["x_opt"])) for a, b in zip(seq, res))
    print(f"batched:    {B} full solves in {dt:.2f} s ({B/dt:.1f} equilibria/s), launches {eng.launches - l0}, {st}, identical to sequential: {same}", flush=True)
    for w in worker_counts:
        st = {}
        t = time.time(); par = qpn_b200.solve_multilevel_workers(net, X, w, engine=eng, stats=st); dt = time.time() - t
        same = all(a["solved"] == b["solved"] and (not a["solved"] or np.array_equal(a["x_opt"], b["x_opt"])) for a, b in zip(res, par))
        print(f"{w:3d} workers: {B} full solves in {dt:.2f} s ({B/dt:.1f} equilibria/s), solved {np.mean([r['solved'] for r in par]):.3f}, {st}, identical to one process: {same}", flush=True)


if __name__ == "__main__":
    main()
