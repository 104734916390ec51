"""CPU suite: the C-ABI library loads and exports every symbol include/qpn_cuda.h declares;
the product path refuses to run without its CUDA library / a GPU (no CPU fallback); the
product-side model and assembly agree with the oracle's independent restatement."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qpn_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qpn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import qpn_b200
    lib = qpn_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libqpn_cuda.so does not export {n}"
    assert set(qpn_b200.engine.EXPORTS) == set(names), set(qpn_b200.engine.EXPORTS) ^ set(names)


def test_no_cpu_fallback():
    import qpn_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(qpn_b200.EngineError, match="no CPU fallback"):
        qpn_b200.Engine(0)
    with pytest.raises(qpn_b200.EngineError):
        qpn_b200.solve(qpn_b200.setup("four_player_matrix_game"))


def test_missing_library_fails_loudly():
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['QPN_CUDA_LIB'] = '/nonexistent/libqpn_cuda.so';"
            "import qpn_b200\ntry:\n    qpn_b200.load_library()\nexcept qpn_b200.EngineError as e:\n    print('LOUD', 'no CPU fallback' in str(e).replace('There is no', 'no'))") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "LOUD True" in out.stdout, out.stdout + out.stderr


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "quadraticprogramnetworks.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r'#include\s*[<"][^>"]*oracle', txt), f        # no native linkage either
                assert "libqpn_oracle" not in txt and "qpo_" not in txt, f


@pytest.mark.parametrize("name", ["simple_bilevel", "four_player_matrix_game", "robust_avoid_simple"])
def test_product_model_equals_oracle_model(name):
    import qpn_b200
    from oracle import examples as oex, qpn_ref
    pn, on = qpn_b200.setup(":" + name), getattr(oex, name)()
    assert pn.n_vars == on.n_vars and pn.network_depth_map == on.depth
    assert pn.network_edges == on.edges and pn.reachable_nodes == on.reach
    for pid in on.qps:
        assert np.array_equal(pn.qps[pid].Q, on.qps[pid]["Q"]) and np.array_equal(pn.qps[pid].q, on.qps[pid]["q"])
        assert pn.qps[pid].var_indices == on.qps[pid]["vars"] and pn.decision_inds(pid) == on.decision_inds(pid)
    for cid, O in qpn_ref.net_polys(on).items():
        P = pn.constraints[cid]
        assert np.array_equal(P.A, O.A) and np.array_equal(P.l, O.l) and np.array_equal(P.u, O.u)
    assert np.array_equal(pn.default_initialization, on.default_init)
    lev = max(on.depth)
    g1, d1, p1 = qpn_b200.assembly.level_gavi(pn, pn.network_depth_map[lev])
    g2, d2, p2 = qpn_ref.level_gavi(on, on.depth[lev], {})
    for k in g2:
        assert g1[k].shape == g2[k].shape and np.array_equal(g1[k], g2[k]), (name, k)
    assert list(d1) == list(d2) and list(p1) == list(p2)
    for pid in on.depth[lev]:
        for a, b in zip(qpn_b200.assembly.node_view(pn, pid), qpn_ref.node_view(on, pid)):
            assert np.array_equal(a, b)


def test_options_mirror_reference_defaults():
    import qpn_b200
    o = qpn_b200.setup("simple_bilevel", gen_solution_map=True).options            # programs.jl:61-77
    assert (o.max_iters, o.num_projections, o.exploration_vertices, o.check_for_cycling, o.gen_solution_map) == (150, 4, 0, True, True)
    o = qpn_b200.setup("robust_avoid_simple").options                              # robust_avoid_simple.jl:4-7,82
    assert (o.exploration_vertices, o.num_projections) == (10, 5)
    with pytest.warns(UserWarning, match="Invalid option name"):
        qpn_b200.setup("simple_bilevel", no_such_option=1)


def test_edge_reduction_and_cycles():
    import qpn_b200
    net = qpn_b200.setup("four_player_matrix_game", edge_list=[(1, 2), (2, 3), (1, 3), (3, 4)])    # 1->3 is redundant
    assert net.network_edges == {1: [2], 2: [3], 3: [4], 4: []}
    assert net.network_depth_map == {1: [1], 2: [2], 3: [3], 4: [4]} and net.reachable_nodes[1] == [2, 3, 4]
    assert net.decision_inds(1) == list(range(8)) and net.decision_inds(4) == [6, 7]
    with pytest.raises(ValueError):
        qpn_b200.setup("four_player_matrix_game", edge_list=[(1, 2), (2, 1)])
