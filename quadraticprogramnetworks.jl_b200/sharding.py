"""Sharding of a batch of independent instances over the GPUs of one box and the final gather.

The path shards by instance: rank g owns the contiguous columns [g*B/G, (g+1)*B/G) of `inits`; there
is no data-path collective.  When a solve is over, every rank packs its results -- x_opt / x_fail, solved flags,
iteration counts, error or pivot counts -- into ONE contiguous block and the blocks are exchanged by ONE collective:
an NCCL all-gather over NVLink on GPUs (gloo in the CPU tests), or, with mode="p2p", peer copies into a symmetric
buffer followed by one symmetric-memory barrier.  One exchange per solve, whatever the number of result arrays.
"""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous range of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BlockLayout:
    """Byte layout of one rank's result block: arrays of fixed dtype / trailing shape, padded to the largest shard."""

    def __init__(self, fields, max_rows):
        # fields: list of (name, numpy dtype, trailing shape)
        self.fields, self.max_rows = [], int(max_rows)
        off = 0
        for name, dtype, tail in fields:
            dtype = np.dtype(dtype)
            row = int(np.prod(tail, dtype=np.int64)) * dtype.itemsize
            off = (off + 15) // 16 * 16
            self.fields.append((name, dtype, tuple(tail), off, row))
            off += row * self.max_rows
        self.nbytes = (off + 15) // 16 * 16

    def pack(self, arrays, buf):
        """arrays: name -> ndarray with first dim <= max_rows; buf: writable uint8 ndarray of nbytes."""
        for name, dtype, tail, off, row in self.fields:
            a = np.ascontiguousarray(arrays[name]).astype(dtype, copy=False)
            n = a.shape[0]
            buf[off: off + n * row] = a.reshape(-1).view(np.uint8)
        return buf

    def unpack(self, buf, rows):
        out = {}
        for name, dtype, tail, off, row in self.fields:
            out[name] = buf[off: off + rows * row].view(dtype).reshape((rows,) + tail)
        return out


class ResultGather:
    """Exchanges one block per rank per solve.  mode "nccl": `all_gather_into_tensor` of the blocks (any backend);
    mode "p2p" (NCCL process groups on one node): every rank copies its block into its slot of every peer's symmetric
    buffer over NVLink and a symmetric-memory barrier publishes them -- no NCCL kernel on the path."""

    def __init__(self, layout, device=None, group=None, mode="nccl"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.layout, self.group = torch, dist, layout, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = device if device is not None else torch.device("cpu")
        self.mode, self.hdl = "nccl", None
        nb = layout.nbytes
        if mode == "p2p" and self.device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm
                self.full = symm.empty(self.world * nb, dtype=torch.uint8, device=self.device)
                self.hdl = symm.rendezvous(self.full, group if group is not None else dist.group.WORLD)
                self.peers = [self.hdl.get_buffer(r, (self.world * nb,), torch.uint8) for r in range(self.world)]
                self.mode = "p2p"
            except Exception:                           # noqa: BLE001 -- symmetric memory is optional
                self.hdl = None
        if self.hdl is None:
            self.full = torch.empty(self.world * nb, dtype=torch.uint8, device=self.device)
        self.block = torch.empty(nb, dtype=torch.uint8, device=self.device)
        pin = self.device.type == "cuda"
        self.host_block = torch.empty(nb, dtype=torch.uint8, pin_memory=pin)
        self.host_full = torch.empty(self.world * nb, dtype=torch.uint8, pin_memory=pin)

    def exchange(self, arrays):
        """Pack this rank's arrays, exchange, return the list of per-rank byte views (host numpy) of all blocks."""
        nb = self.layout.nbytes
        self.layout.pack(arrays, self.host_block.numpy())
        self.block.copy_(self.host_block, non_blocking=True)
        self.exchange_device()
        self.host_full.copy_(self.full, non_blocking=False)
        hf = self.host_full.numpy()
        return [hf[r * nb: (r + 1) * nb] for r in range(self.world)]

    def exchange_device(self):
        """`self.block` (device) -> `self.full` on every rank."""
        nb = self.layout.nbytes
        if self.mode == "p2p":
            for r in range(self.world):
                self.peers[r][self.rank * nb: (self.rank + 1) * nb].copy_(self.block, non_blocking=True)
            self.hdl.barrier(channel=0)
        else:
            self.dist.all_gather_into_tensor(self.full, self.block, group=self.group)


def gather_results(local, batch, group=None, device=None, mode="nccl"):
    """local: name -> numpy array (first dim = this rank's shard).  Every rank gets the full-batch arrays; ONE collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    layout = BlockLayout([(k, v.dtype, v.shape[1:]) for k, v in local.items()], max(hi - lo for lo, hi in sizes))
    g = ResultGather(layout, device=device, group=group, mode=mode)
    blocks = g.exchange(local)
    parts = [layout.unpack(b, hi - lo) for b, (lo, hi) in zip(blocks, sizes)]
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in local}


def _device_for(group):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def solve_sharded(solver, inits, group=None, mode="nccl"):
    """solve(qpn, inits) of a flat game with the batch sharded over the ranks of `group`; every rank gets all results."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(inits), rank, world)
    ret = solver.solve_batch(np.ascontiguousarray(inits[lo:hi]))
    local = dict(x=np.asarray(ret["x"], np.float64), solved=np.asarray(ret["solved"], np.uint8),
                 iters=np.asarray(ret["iters"], np.int32), pivots=np.asarray(ret["pivots"], np.int32))
    return gather_results(local, len(inits), group, device=_device_for(group), mode=mode)


def solve_net_sharded(binding, inits, group=None, mode="nccl"):
    """solve(qpn, inits) of a network with children (netsolve.NetBinding) sharded over the ranks of `group`: each rank
    runs the native state machine on its columns, one exchange brings x / solved / level_iters / error to every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(len(inits), rank, world)
    ret = binding.solve_arrays(np.ascontiguousarray(inits[lo:hi]))
    local = dict(x=ret["x"], solved=np.asarray(ret["solved"], np.uint8), level_iters=np.asarray(ret["level_iters"], np.int32),
                 error=np.asarray(ret["error"], np.int32))
    out = gather_results(local, len(inits), group, device=_device_for(group), mode=mode)
    out["solved"] = out["solved"].astype(bool)
    return out
