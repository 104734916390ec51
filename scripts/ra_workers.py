"""Full three-level robust_avoid_simple batches (BASELINE.json configs[2]) through a MultilevelPool: host worker
processes served by one engine handle.  usage: ra_workers.py B workers [workers ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200


def main():
    B = int(sys.argv[1])
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    rng = np.random.default_rng(0)
    X = np.tile(net.default_initialization, (B, 1)); X[:, 0:6] += 0.3 * rng.normal(size=(B, 6)); X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    eng = qpn_b200.Engine(0)
    for w in [int(a) for a in sys.argv[2:]]:
        t = time.time()
        with qpn_b200.MultilevelPool(net, w, engine=eng, chunk=int(os.environ.get('QPN_CHUNK', '256'))) as pool:
            pool.solve(X[:4 * w]); t_up = time.time() - t                     # processes up, memos warm
            st = {}
            t = time.time(); par = pool.solve(X, stats=st); dt = time.time() - t
        print(f"{w:3d} workers: {B} in {dt:.2f} s ({B/dt:.1f} equilibria/s), pool up in {t_up:.1f} s, solved {np.mean([r['solved'] for r in par]):.4f}, "
              f"{ {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()} }", flush=True)


if __name__ == "__main__":
    main()
