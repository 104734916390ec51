"""CPU suite: the N>1 path (contiguous sharding + one final all-gather) under gloo, world_size 2.
The numeric stand-in on the CPU is the oracle's level loop; on GPUs the same host logic drives
libqpn_cuda with the NCCL backend (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_batch():
    import qpn_b200
    for B in (0, 1, 7, 4096, 65536, 65537):
        for W in (1, 2, 3, 8):
            r = [qpn_b200.sharding.shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import qpn_b200
    from oracle import cport, examples, qpn_ref
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    L = cport.Level(8, [qpn_ref.node_view(net, p) for p in net.depth[1]], g, dec, par, 150, None)

    class CpuStandIn:                                # same interface as BatchedSolver.solve_batch
        def solve_batch(self, inits):
            return L.solve(inits)
    X = np.random.default_rng(99).uniform(-5, 5, (B, 8))
    full = qpn_b200.sharding.solve_sharded(CpuStandIn(), X)
    if rank == 0:
        q.put({k: v for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_equals_unsharded_gloo():
    import torch.multiprocessing as mp
    from oracle import cport, examples, qpn_ref
    B, world = 101, 2                                     # odd: the shards differ in size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    L = cport.Level(8, [qpn_ref.node_view(net, p) for p in net.depth[1]], g, dec, par, 150, None)
    X = np.random.default_rng(99).uniform(-5, 5, (B, 8))
    ref = L.solve(X)
    assert np.array_equal(full["x"], ref["x"]) and np.array_equal(full["pivots"], ref["pivots"])
    assert np.array_equal(full["solved"].astype(bool), ref["solved"]) and np.array_equal(full["iters"], ref["iters"])


def _net_worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import qpn_b200
    from tests.native_oracle import oracle_net
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    X = qpn_b200.examples.robust_avoid_batch(net, B, seed=5)
    full = qpn_b200.sharding.solve_net_sharded(oracle_net(net, threads=2), X)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_network_solve_equals_unsharded_gloo():
    """solve(qpn, inits) of the three-level robust_avoid net sharded over two ranks (native state machine on the oracle
    numerics per rank, ONE exchange of the packed result block) == the unsharded batch."""
    import torch.multiprocessing as mp
    import qpn_b200
    from tests.native_oracle import oracle_net
    B, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_net_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    net = qpn_b200.setup("robust_avoid_simple", seed=3)
    ref = oracle_net(net).solve_arrays(qpn_b200.examples.robust_avoid_batch(net, B, seed=5))
    for k in ("x", "solved", "level_iters", "error"):
        assert np.array_equal(full[k], ref[k]), k
