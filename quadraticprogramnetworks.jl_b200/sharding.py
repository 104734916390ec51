"""Sharding of a batch of independent instances over the GPUs of one box and the final gather.

The path shards by instance: rank g owns the contiguous columns
[g*B/G, (g+1)*B/G) of `inits`; there is no data-path collective, only one all-gather of
x_opt / solved / iters / pivots at the end (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous range of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, batch, group=None):
    """All-gather dict of per-instance torch tensors (first dim = local batch) into full-batch
    tensors on every rank.  Shards may differ in size by one, so shorter ones are padded."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    out = {}
    for key, t in local.items():
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        full = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, pad, group=group)
        out[key] = torch.cat([full[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return out


def solve_sharded(solver, inits, group=None):
    """solve(qpn, inits) with the batch sharded over the ranks of `group`; every rank gets all results."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(inits), rank, world)
    ret = solver.solve_batch(np.ascontiguousarray(inits[lo:hi]))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    local = {k: torch.from_numpy(np.ascontiguousarray(ret[k])).to(dev) for k in ("x", "solved", "iters", "pivots")}
    local["solved"] = local["solved"].to(torch.uint8)
    full = gather_results(local, len(inits), group)
    return {k: v.cpu().numpy() for k, v in full.items()}
