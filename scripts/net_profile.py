"""Per-kernel accounting of one full three-level robust_avoid batch through the native state machine (option "profile":
every launch bracketed by CUDA events on its stream).  usage: net_profile.py [B] [threads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qpn_b200
from qpn_b200.netsolve import NetBinding
from tests.native_oracle import ra_inits


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    X = ra_inits(net, B, seed=0)
    eng = qpn_b200.Engine(0)
    nb = NetBinding(net, eng.lib, "qpn_net_", handle=eng.h, threads=t)
    nb.solve_arrays(X)
    nb.solve_arrays(X)
    nb.set_option("profile", 1)
    p0, s0 = nb.profile(), nb.stats()
    t0 = time.time(); nb.solve_arrays(X); dt = time.time() - t0
    p1, s1 = nb.profile(), nb.stats()
    print(f"B={B} threads={t}: {dt:.3f}s wall (profile mode), rounds {s1['rounds'] - s0['rounds']}, host {(s1['host_ns'] - s0['host_ns']) / 1e6:.1f} ms, "
          f"backend {(s1['backend_ns'] - s0['backend_ns']) / 1e6:.1f} ms")
    for k in ("verify", "solve_qep", "member", "group", "cycle"):
        d = {f: p1[k][f] - p0[k][f] for f in ("launches", "units", "ms")}
        print(f"  {k:10s} launches {d['launches']:6d}  units {d['units']:10d}  ms {d['ms']:9.2f}  ms/launch {d['ms'] / max(d['launches'], 1):.4f}  "
              f"us/unit {1e3 * d['ms'] / max(d['units'], 1):.3f}")
    print(f"  h2d {p1['h2d_bytes'] - p0['h2d_bytes']} B, d2h {p1['d2h_bytes'] - p0['d2h_bytes']} B")


if __name__ == "__main__":
    main()
