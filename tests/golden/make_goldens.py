"""Regenerates the fixtures in this directory.

  simple_bilevel_kat.json   hand-transcribed from /root/reference/test/simple_bilevel.jl:4-16 (the
                            reference's only pinned results) plus KAT-0 of SURVEY.md 8c;
  oracle_*.json             outputs of the C oracle (oracle/qpn_oracle.c) on seeded inputs.  The
                            true reference (Julia + PATH + OSQP) cannot run in this container, so
                            these pin the oracle against regressions; they are NOT reference outputs.

Run from the repo root:  python tests/golden/make_goldens.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cport, examples, qpn_ref  # noqa: E402
from tests import problems  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)


def main():
    s2 = math.sqrt(2.0)
    dump("simple_bilevel_kat.json", {
        "source": "/root/reference/test/simple_bilevel.jl:4-21 (x_opt = [w; x; y], atol 1e-4; min piece counts S)",
        "W": [[-2.0, -3.0], [0.0, -1.0], [1.0, -3.0], [1.0, -1.0], [1.0, 0.0], [0.0, 1.0], [-1.0, 1 + s2], [0.0, 0.0]],
        "X": [[[-2.0, 0.0]], [[0.0, 0.0]], [[0.0, 0.0]], [[0.0, 0.0]], [[0.5, 0.5]], [[0.5, 0.5], [0.0, 0.0]],
              [[-1.0, 0.0], [s2 / 2, s2 / 2]], [[0.0, 0.0]]],
        "S": [1, 2, 1, 2, 1, 1, 1, 3],
        "kat0": {"source": "SURVEY.md 8c: level-2 AVI of simple_bilevel, z = [y; xi; lambda; s], w = [w1, w2, x]",
                 "M": [[0, 1, 0, 0], [2, 0, -1, 0], [1, 0, 0, -1], [0, 0, 1, 0]],
                 "N": [[0, 0, 0], [0, 0, -2], [0, 0, 0], [0, 0, 0]],
                 "l": ["-inf", "-inf", "-inf", 0], "u": ["inf", "inf", "inf", "inf"],
                 "solution": "z = [max(x,0), 0, max(-2x,0), max(x,0)]"}})
    rng = np.random.default_rng(20261018)
    net, g, avi, dec, par = problems.fp_avi()
    X, z0 = problems.fp_starts(rng, 12)
    z, st, pv, bs = cport.avi_solve_batched(avi["M"], np.tile(avi["o"], (12, 1)), avi["l"], avi["u"], z0)
    dump("oracle_four_player_avi.json", {"inits": X.tolist(), "z": z.tolist(), "status": st.tolist(), "pivots": pv.tolist(), "basis": bs.tolist()})
    net, X = problems.ra_inits(rng, 8)
    L = cport.Level(net.n_vars, [qpn_ref.node_view(net, p) for p in net.depth[3]], *qpn_ref.level_gavi(net, net.depth[3], {}), 150, None)
    r = L.solve(X)
    dump("oracle_robust_avoid_bottom_level.json", {"inits": X.tolist(), "x": r["x"].tolist(), "solved": r["solved"].tolist(),
                                                   "iters": r["iters"].tolist(), "pivots": r["pivots"].tolist(), "lam": r["lam"].tolist()})


if __name__ == "__main__":
    main()
