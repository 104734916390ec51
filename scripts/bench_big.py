"""Config 5 of BASELINE.json (synthetic monotone AVI stress test n=256, m=512) through
the global-memory tableau path: timing, pivots, KKT residuals, and parity of a few instances against
the C oracle.  usage: python scripts/bench_big.py [n m batch [n_oracle]]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200  # noqa: E402
from oracle import cport  # noqa: E402
from tests.test_gpu_big import monotone_gavi  # noqa: E402


def main():
    n, m, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 512, 296)
    n_oracle = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rng = np.random.default_rng(5)
    g, xbar = monotone_gavi(rng, n, m)
    g["N"] = np.eye(n); g["B"] = np.zeros((m, n))
    O = rng.normal(size=(B, n))
    z0 = np.zeros((B, n + m)); z0[:, :n] = xbar
    eng = qpn_b200.Engine(0)
    ga = qpn_b200.engine.GaviArrays(g)
    for rep in range(2):
        t = time.time()
        ret = eng.gavi_solve(ga, O, z0)
        dt = time.time() - t
        print(f"n={n} m={m} lifted={n + 2 * m} B={B}: {dt * 1e3:.1f} ms  ({B / dt:.1f} solves/s)  pivots p50={int(np.median(ret['pivots']))} "
              f"ok={(ret['status'] == 1).all()}", flush=True)
    x = ret["z"][:, :n]; lam = ret["z"][:, n:]
    Q, A = g["M"][:, :n], g["A"][:, :n]
    print("stationarity", np.abs(x @ Q.T + O - lam @ A).max(), "feas", (x @ A.T - g["l2"]).min(), "lam min", lam.min(),
          "compl", np.abs(lam * (x @ A.T - g["l2"])).max())
    piv = ret["pivots"].astype(np.float64)
    print("pivots total", piv.sum(), "instance-pivots/s", piv.sum() / dt)
    for k in range(n_oracle):
        t = time.time()
        ro = cport.gavi_solve(g, z0[k], O[k])
        print(f"oracle[{k}]: {time.time() - t:.2f} s status {ro['status']} pivots {ro['pivots']} ; GPU pivots {ret['pivots'][k]} "
              f"basis equal {np.array_equal(ro['basis'], ret['basis'][k])} z bit-equal {np.array_equal(ro['z_full'], ret['z_full'][k])}")


if __name__ == "__main__":
    main()
