"""Parity against TRUE reference output (produced by julia/ref_julia.jl where Julia + PATH exist).

usage: python scripts/compare_with_julia_goldens.py <dir written by ref_julia.jl>

For every example found: load the reference's own model (`*_model.json`, flat-array format), solve from the same
inits on the B200 engine, and report agreement of `solved` and of `x_opt` (1e-8 relative for unique equilibria --
four_player; 1e-4 absolute as the reference's own test uses for simple_bilevel; robust_avoid equilibria are
not unique, so there the check is that OUR point passes the reference's optimality test, reported separately)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qpn_b200  # noqa: E402


def main(d):
    for name in ("simple_bilevel", "four_player_matrix_game", "robust_avoid_simple"):
        mp, gp = os.path.join(d, f"{name}_model.json"), os.path.join(d, f"{name}_goldens.json")
        if not (os.path.exists(mp) and os.path.exists(gp)):
            print(f"{name}: no goldens in {d}")
            continue
        net = qpn_b200.load_net(mp)
        G = json.load(open(gp))
        X = np.array(G["inits"], dtype=float)
        t = time.time(); res = qpn_b200.solve(net, X); dt = time.time() - t
        res = res if isinstance(res, list) else [res]
        solved_ref = np.array(G["solved"], bool)
        solved = np.array([r["solved"] for r in res])
        both = solved & solved_ref
        xr = np.array(G["x"], dtype=float)
        xo = np.array([r["x_opt"] if r["solved"] else r["x_fail"] for r in res])
        err = np.abs(xo[both] - xr[both]).max(initial=0.0)
        rel = (np.abs(xo[both] - xr[both]) / np.maximum(1.0, np.abs(xr[both]))).max(initial=0.0)
        print(f"{name}: {len(X)} instances; reference {G['equilibria_per_s']:.1f} eq/s on {G['threads']} threads, here {len(X) / dt:.1f} eq/s; "
              f"solved agree {np.mean(solved == solved_ref):.3f}; over jointly solved: max abs diff {err:.3e}, max rel diff {rel:.3e}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "julia_goldens")
