"""Multi-level batches over several host processes and ONE engine handle.

With the device calls of a batch regrouped (batching.py), what bounds a batch of networks with children is the host
side of `solve_base!` (/root/reference/src/algorithm.jl:1-127): piece generation, set operations, cycle detection --
independent per instance and, in CPython, serialised by the interpreter lock within a process.  A `MultilevelPool`
therefore keeps `workers` host processes alive.  Each runs the unchanged batched state machine
(`solve_multilevel_batch`) on a contiguous shard of the instances (the rule of sharding.shard_range, as across GPUs)
against a `RemoteEngine`: a proxy that forwards every regrouped device call over a pipe to the process that owns the
GPU.  There is one engine handle (one CUDA context, one set of device buffers) however many workers there are;
problem data of a call (a node's matrices, a GAVI) crosses the pipe once per worker and is referred to by a small
integer afterwards.  The workers' piece / predicate memos live as long as the pool, like a resident level.

Every instance is solved by the same code against the same device engine as in one process, so the results are
identical (tests/test_multilevel_cpu.py, tests/test_gpu_parity.py).
"""
import multiprocessing as mp
import threading
import time

import numpy as np

from .sharding import shard_range


class RemoteEngine:
    """The engine as a worker process sees it: `remote_call` is what BatchingEngine._execute hands a regrouped
    device call to.  Only that form exists -- a worker never talks to the device any other way."""

    def __init__(self, conn):
        self.conn = conn
        self.launches = 0
        self._ids = {}

    def remote_call(self, kind, key, shared, stacked, kwargs):
        ident = self._ids.get((kind, key))
        first = ident is None
        if first:
            ident = self._ids[(kind, key)] = len(self._ids)
        self.conn.send(("call", kind, ident, shared if first else None, stacked, kwargs))
        tag, payload = self.conn.recv()
        if tag == "err":
            raise RuntimeError(payload)
        self.launches += 1
        return payload


def _pool_worker(conn, qpn, chunk):
    """Worker process: waits for shards, solves them, sends the results back; memos persist between shards."""
    from .algorithm import solve_multilevel_batch
    eng = RemoteEngine(conn)
    pieces, memo = {}, {}
    while True:
        msg = conn.recv()
        if msg[0] == "stop":
            break
        _, X, keep_sol = msg
        stats = {}
        t0 = time.perf_counter()
        try:
            outs = solve_multilevel_batch(qpn, X, eng, chunk=chunk, stats=stats, pieces=pieces, memo=memo)
            if not keep_sol:
                for r in outs:
                    r.pop("Sol", None)
            stats["solve_s"] = time.perf_counter() - t0
            conn.send(("done", outs, stats))
        except BaseException as e:                       # noqa: BLE001 -- reported to the caller of pool.solve
            conn.send(("fail", f"{type(e).__name__}: {e}", None))
    conn.close()


class MultilevelPool:
    """`workers` host processes for the multi-level batches of one QPNet, all served by one engine."""

    def __init__(self, qpn, workers, engine=None, device=0, chunk=256):
        from .engine import Engine
        self.engine = engine if engine is not None else Engine(device)
        self._device_engine = isinstance(self.engine, Engine)
        self.workers = max(1, int(workers))
        self.lock = threading.Lock()                     # calls on a handle are serialised by the caller (include/qpn_cuda.h)
        self.device_calls = 0
        ctx = mp.get_context("spawn")                    # spawn: the children must not inherit the CUDA context
        self.conns, self.procs = [], []
        for _ in range(self.workers):
            here, there = ctx.Pipe()
            p = ctx.Process(target=_pool_worker, args=(there, qpn, chunk), daemon=True)
            p.start()
            there.close()
            self.conns.append(here); self.procs.append(p)
        self._shared = [dict() for _ in range(self.workers)]     # per worker: ident -> marshalled problem data

    def _marshal(self, kind, shared):
        if not self._device_engine:
            return shared
        from .engine import GaviArrays, NodeArrays
        if kind == "verify":
            return NodeArrays(*shared)
        if kind in ("gavi", "comp"):
            return GaviArrays(shared)
        return shared

    def _execute(self, w, kind, ident, shared, stacked, kwargs):
        table = self._shared[w]
        if shared is not None:
            table[ident] = self._marshal(kind, shared)
        data = table[ident]
        with self.lock:
            self.device_calls += 1
            if kind == "verify":
                return self.engine.verify_solution(data, stacked[0], **kwargs)
            if kind == "gavi":
                return self.engine.gavi_solve(data, stacked[0], stacked[1], **kwargs)
            if kind == "comp":
                return self.engine.comp_indices(data, stacked[0], stacked[1], **kwargs)
            return self.engine.halfspace_in(data, stacked[0], **kwargs)

    def _serve(self, w, X, keep_sol, out):
        conn = self.conns[w]
        try:
            conn.send(("solve", X, keep_sol))
            while True:
                msg = conn.recv()
                if msg[0] == "done":
                    out[w] = (msg[1], msg[2]); return
                if msg[0] == "fail":
                    out[w] = RuntimeError(f"worker {w}: {msg[1]}"); return
                try:
                    conn.send(("ret", self._execute(w, *msg[1:])))
                except Exception as e:                   # noqa: BLE001 -- delivered to the instances that asked
                    conn.send(("err", f"{type(e).__name__}: {e}"))
        except (EOFError, OSError) as e:
            out[w] = RuntimeError(f"worker {w} died: {e}")

    def solve(self, X, keep_sol=False, stats=None):
        """solve(qpn, inits) for the whole batch; results in instance order.  `Sol` (the top level's solution
        pieces) stays in the workers unless keep_sol: it is large and a batch caller reads x_opt / solved."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        used = min(self.workers, len(X))
        out = [None] * used
        calls0 = self.device_calls
        threads = [threading.Thread(target=self._serve, args=(w, X[slice(*shard_range(len(X), w, used))], keep_sol, out), daemon=True)
                   for w in range(used)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for o in out:
            if isinstance(o, Exception):
                raise o
        if stats is not None:
            for _, st in out:
                for k, v in st.items():
                    stats[k] = stats.get(k, 0) + v
            stats["workers"] = used
            stats["engine_calls"] = self.device_calls - calls0
        return [r for part, _ in out for r in part]

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except (OSError, BrokenPipeError):
                pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
        for c in self.conns:
            c.close()
        self.conns, self.procs = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def solve_multilevel_workers(qpn, X, workers, engine=None, device=0, chunk=256, stats=None, keep_sol=False):
    """One-shot form: a pool for this batch only."""
    with MultilevelPool(qpn, min(int(workers), len(X)), engine=engine, device=device, chunk=chunk) as pool:
        return pool.solve(X, keep_sol=keep_sol, stats=stats)
