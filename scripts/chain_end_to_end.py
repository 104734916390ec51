"""BASELINE configs[3] end to end: the synthetic 3-level chain through the native state machine (solve(qpn, inits)),
for growing numbers of variables per node.  usage: chain_end_to_end.py [oracle|device] [B] [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
from qpn_b200.netsolve import NetBinding


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "oracle"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ns = [int(a) for a in sys.argv[3:]] or [2, 4, 6, 8]
    eng = qpn_b200.Engine(0) if backend == "device" else None
    for n in ns:
        net = qpn_b200.setup("synthetic_chain", n=n, levels=3, n_params=8)
        rng = np.random.default_rng(n)
        X = np.tile(net.default_initialization, (B, 1)) + rng.normal(size=(B, net.n_vars))
        if backend == "device":
            nb = NetBinding(net, eng.lib, "qpn_net_", handle=eng.h, threads=2)
        else:
            from tests.native_oracle import oracle_net
            nb = oracle_net(net, threads=os.cpu_count())
        t0 = time.time(); r = nb.solve_arrays(X); cold = time.time() - t0
        s0 = nb.stats()
        t0 = time.time(); r = nb.solve_arrays(X); warm = time.time() - t0
        s1 = nb.stats()
        print(f"n={n} per node ({net.n_vars} vars): cold {cold:.2f}s (LPs {s0['lps']}, pieces {s0['pieces']}), warm {B} instances in {warm:.3f}s = "
              f"{B / warm:,.0f}/s, solved {r['solved'].mean():.3f}, mean passes per level {r['level_iters'].mean(0).round(2).tolist()}, "
              f"errors {np.bincount(r['error'] & 0xff).tolist()}", flush=True)
        nb.close()


if __name__ == "__main__":
    main()
