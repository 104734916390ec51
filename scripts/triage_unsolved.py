"""Triage of the unsolved instances of the full three-level robust_avoid batch: exit reason of
every instance that does not reach an equilibrium, by level and by the StatusCode the failing solve_qep returned.
Runs the oracle build of the native state machine on all host cores, or -- with QPN_TRIAGE_DEVICE=1 -- the device (same
results: tests/test_gpu_net.py).
usage: triage_unsolved.py [B] [data seeds ...]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
from qpn_b200.netsolve import ERRORS
from tests.native_oracle import oracle_net

STATUS = {1: "SUCCESS", 2: "RAY_TERM", 3: "MAX_ITERS", 4: "FAILURE (final check_avi_solution)"}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    seeds = [int(a) for a in sys.argv[2:]] or [3]
    print("| data seed | instances | solved | exit reason | level | solve_qep status | count | share |")
    print("|---|---|---|---|---|---|---|---|")
    for seed in seeds:
        net = qpn_b200.setup("robust_avoid_simple", seed=seed)
        X = qpn_b200.examples.robust_avoid_batch(net, B, seed=0)
        if os.environ.get("QPN_TRIAGE_DEVICE") == "1":
            from qpn_b200.netsolve import NetBinding
            eng = qpn_b200.Engine(0)
            r = NetBinding(net, eng.lib, "qpn_net_", handle=eng.h, threads=2).solve_arrays(X)
        else:
            r = oracle_net(net, threads=os.cpu_count()).solve_arrays(X)
        c = collections.Counter(int(e) for e in r["error"][~r["solved"]])
        for code, n in sorted(c.items(), key=lambda kv: -kv[1]):
            low, st, lv = code & 0xff, (code >> 8) & 0xff, (code >> 16) & 0xff
            print(f"| {seed} | {B} | {r['solved'].mean():.4f} | {ERRORS[low][:60]} | {lv + 1 if low == 2 else '-'} | "
                  f"{STATUS.get(st, '-') if low == 2 else '-'} | {n} | {n / B:.5f} |")
        if not c:
            print(f"| {seed} | {B} | 1.0000 | - | - | - | 0 | 0 |")


if __name__ == "__main__":
    main()
