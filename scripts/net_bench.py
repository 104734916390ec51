"""Full three-level robust_avoid_simple batches (BASELINE.json configs[2]) through the native state machine:
device (qpn_net_*) timing for several host thread counts, against the oracle build of the same logic.
usage: net_bench.py [B] [threads ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
from qpn_b200.netsolve import NetBinding
from tests.native_oracle import oracle_net, ra_inits


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    threads = [int(a) for a in sys.argv[2:]] or [1, 4, 8]
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    X = ra_inits(net, B, seed=0)
    eng = qpn_b200.Engine(0)
    ref = None
    for t in threads:
        nb = NetBinding(net, eng.lib, "qpn_net_", handle=eng.h, threads=t)
        t0 = time.time(); nb.solve_arrays(X); cold = time.time() - t0
        s0 = nb.stats()
        t0 = time.time(); ret = nb.solve_arrays(X); dt = time.time() - t0
        s1 = nb.stats()
        d = {k: s1[k] - s0[k] for k in s1}
        print(f"device threads={t}: cold {cold:.3f}s; warm {B} equilibria in {dt:.3f}s = {B / dt:,.0f}/s, solved {ret['solved'].mean():.4f}, "
              f"launches {d['launches']}, rounds {d['rounds']}, calls {d['calls']}, requests {d['requests']}, new lps {d['lps']}, "
              f"host {d['host_ns'] / 1e6:.1f} ms, backend {d['backend_ns'] / 1e6:.1f} ms (summed over threads), cohort splits {d['cohort_splits']}", flush=True)
        if ref is None:
            ref = ret
        else:
            print("   identical to first run:", all(np.array_equal(ref[k], ret[k]) for k in ref))
        nb.close()
    nc = os.cpu_count()
    ob = oracle_net(net, threads=nc)
    n_or = min(B, 4096)
    ob.solve_arrays(X[:256])
    t0 = time.time(); ro = ob.solve_arrays(X[:n_or]); dt = time.time() - t0
    print(f"oracle build, {nc} threads: {n_or} equilibria in {dt:.3f}s = {n_or / dt:,.0f}/s")
    print("device == oracle on those:", all(np.array_equal(ref[k][:n_or], ro[k]) for k in ro))


if __name__ == "__main__":
    main()
