"""CPU suite for the native network state machine (csrc/net/): the oracle-backed build of
the same C++ host logic against the Python host mirror (`NetSolver`) on the same oracle numerics -- status, x,
per-level iteration counts and the solution graphs must agree bit for bit -- plus the reference's own known answers
(test/simple_bilevel.jl:4-21) through the native path."""
import json
import os

import numpy as np
import pytest

import qpn_b200
from tests.native_oracle import oracle_net, ra_inits, same_result
from tests.oracle_engine import OracleEngine

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mirror_results(net, X):
    eng, pieces, memo = OracleEngine(), {}, {}
    return [qpn_b200.NetSolver(net, eng, piece_cache=pieces, lp_memo=memo).solve(x) for x in X]


def test_native_simple_bilevel_known_answers():
    """KAT-1..8 (test/simple_bilevel.jl:4-21) through qpo_net_solve_batched, all eight as ONE batch."""
    kat = json.load(open(os.path.join(GOLDEN, "simple_bilevel_kat.json")))
    net = qpn_b200.setup(":simple_bilevel", gen_solution_map=True)
    nb = oracle_net(net)
    X = np.array([w + [0.0, 0.0] for w in kat["W"]])
    outs = nb.solve(X, keep_sol=True)
    for w, Xs, s, ret in zip(kat["W"], kat["X"], kat["S"], outs):
        assert ret["solved"], ret.get("error")
        assert any(np.allclose(ret["x_opt"], w + xi, atol=1e-4) for xi in Xs), (w, ret["x_opt"])     # :19
        assert len(ret["Sol"][2]) >= s                                                             # :20
    ref = mirror_results(net, X)
    assert all(same_result(a, b, sol=True) for a, b in zip(outs, ref))


@pytest.mark.parametrize("seed,B", [(3, 64), (5, 48), (1, 32)])
def test_native_matches_mirror_on_robust_avoid(seed, B):
    """examples/robust_avoid_simple.jl, three levels: native == mirror for every instance (seed 1 ends in the
    reference's cycling exit for many instances: the failure path and x_fail are compared too)."""
    net = qpn_b200.setup("robust_avoid_simple", seed=seed)
    X = ra_inits(net, B, seed=seed)
    nb = oracle_net(net)
    nat = nb.solve(X, keep_sol=True)
    ref = mirror_results(net, X)
    bad = [b for b, (a, r) in enumerate(zip(nat, ref)) if not same_result(a, r, sol=True)]
    assert not bad, bad
    st = nb.stats()
    assert st["calls"] < st["requests"] / 4                 # requests really are regrouped across instances
    # a second batch reuses every piece: no new geometry, identical answers
    again = nb.solve(X)
    st2 = nb.stats()
    assert st2["lps"] == st["lps"] and st2["pieces"] == st["pieces"]
    assert all(same_result(a, r) for a, r in zip(again, ref))


def test_native_threads_and_batch_composition_do_not_change_results():
    """Host threads share the memo of exact geometry: any thread count, any batch split, same bits."""
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    X = ra_inits(net, 45, seed=11)
    one = oracle_net(net, threads=1).solve(X)
    many = oracle_net(net, threads=4).solve(X)
    assert all(same_result(a, b) for a, b in zip(one, many))
    nb = oracle_net(net, threads=2)
    split = nb.solve(X[30:]) + nb.solve(X[:30])
    assert all(same_result(a, b) for a, b in zip(one[30:] + one[:30], split))


def test_native_reports_failures_per_instance():
    """algorithm.jl:120-126: an instance that cannot be solved is reported, the batch goes on."""
    net = qpn_b200.setup(":simple_bilevel")
    net.options.max_iters = 1
    nb = oracle_net(net)
    outs = nb.solve(np.array([[1.0, 0.0, 3.0, 0.0], [1.0, 0.0, 0.0, 0.0]]))
    assert (not outs[0]["solved"]) and outs[0]["x_opt"] is None and "x_fail" in outs[0] and "Can't find" in outs[0]["error"]
    ref = mirror_results(net, np.array([[1.0, 0.0, 3.0, 0.0], [1.0, 0.0, 0.0, 0.0]]))
    assert all(same_result(a, b) for a, b in zip(outs, ref))


@pytest.mark.parametrize("edges", [[(1, 2), (3, 4)], [(1, 2), (2, 3), (3, 4)], [(1, 2), (1, 3), (1, 4)], [(1, 2), (2, 3)]])
def test_native_matches_mirror_on_hierarchical_four_player(edges):
    """examples/four_player_matrix_game.jl with edges (the example sweeps the power set of edge lists: parallel bilevel,
    a four-level chain, one leader with three followers, ...): native == mirror, every instance solved."""
    net = qpn_b200.setup("four_player_matrix_game", edge_list=edges)
    X = np.random.default_rng(7).uniform(-5.0, 5.0, (24, 8))
    nat = oracle_net(net, threads=2).solve(X, keep_sol=True)
    ref = mirror_results(net, X)
    assert all(r["solved"] for r in nat)
    bad = [b for b, (a, r) in enumerate(zip(nat, ref)) if not same_result(a, r, sol=True)]
    assert not bad, bad


def test_wide_batched_subset_lps_change_no_result():
    """The CUDA backend batches the subset tests of remove_subsets across all pairs of a list (LPBackend::wide_batches); the
    oracle build takes the same path with QPN_ORACLE_WIDE=1 (read once per process): same results, fewer LP calls."""
    import subprocess, sys
    code = ("import sys, json, hashlib; sys.path.insert(0, %r); import numpy as np, qpn_b200\n"
            "from tests.native_oracle import oracle_net, ra_inits\n"
            "net = qpn_b200.setup('robust_avoid_simple', seed=3); X = ra_inits(net, 160, seed=3)\n"
            "nb = oracle_net(net, threads=2); r = nb.solve_arrays(X); s = nb.stats()\n"
            "h = hashlib.sha256(r['x'].tobytes() + r['solved'].tobytes() + r['level_iters'].tobytes()).hexdigest()\n"
            "print(json.dumps(dict(h=h, calls=s['lp_calls'], lps=s['lps'])))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for wide in (False, True):
        env = dict(os.environ)
        env.pop("QPN_ORACLE_WIDE", None)
        if wide:
            env["QPN_ORACLE_WIDE"] = "1"
        outs.append(json.loads(subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip().splitlines()[-1]))
    assert outs[0]["h"] == outs[1]["h"]
    assert outs[1]["calls"] < outs[0]["calls"]
