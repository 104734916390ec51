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

Calls are merged across workers: one dispatcher thread owns the engine; while a device call runs, the calls the other
workers send queue up, and the dispatcher then takes ALL of them, groups those with the same (kind, problem data,
options) -- the workers walk the same network, so the popular ones coincide -- and issues one batched device call
per group.  Nothing waits for a merge partner: an idle engine serves a lone call at once.

Every instance is solved by the same code against the same device engine as in one process, so the results are
identical (tests/test_multilevel_cpu.py, tests/test_gpu_parity.py).
"""
import multiprocessing as mp
import threading
import time

import numpy as np

from .engine import EngineError

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
            raise EngineError(payload)                     # the device call failed in the process that owns the GPU
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


class _Pending:
    __slots__ = ("kind", "gid", "stacked", "kwargs", "result", "error", "ready")

    def __init__(self, kind, gid, stacked, kwargs):
        self.kind, self.gid, self.stacked, self.kwargs = kind, gid, stacked, kwargs
        self.result, self.error, self.ready = None, None, threading.Event()


class MultilevelPool:
    """`workers` host processes for the multi-level batches of one QPNet, all served by one engine."""

    def __init__(self, qpn, workers, engine=None, device=0, chunk=256):
        from .engine import Engine
        self.engine = engine if engine is not None else Engine(device)
        self._device_engine = isinstance(self.engine, Engine)
        self.workers = max(1, int(workers))
        self.broken = False
        self.device_calls = 0            # engine calls issued (after merging)
        self.worker_calls = 0            # calls received from the workers
        ctx = mp.get_context("spawn")                    # spawn: the children must not inherit the CUDA context
        self.conns, self.procs = [], []
        for _ in range(self.workers):
            here, there = ctx.Pipe()
            p = ctx.Process(target=_pool_worker, args=(there, qpn, chunk), daemon=True)
            p.start()
            there.close()
            self.conns.append(here); self.procs.append(p)
        self._ident = [dict() for _ in range(self.workers)]      # per worker: its ident -> global id of the problem data
        self._gid, self._data = {}, []                           # bytes key -> global id ; global id -> marshalled data
        self._meta = threading.Lock()
        self._queue, self._cv, self._stop = [], threading.Condition(), False
        self._dispatcher = threading.Thread(target=self._dispatch, daemon=True)
        self._dispatcher.start()

    # ---- problem data: crosses the pipe once per worker, marshalled once per pool --------------------------
    def _register(self, w, kind, ident, shared):
        from .batching import _gavi_key, _key
        if kind == "verify":
            key = ("n", _key(*shared))
        elif kind in ("gavi", "comp"):
            key = ("g", _gavi_key(shared))
        else:
            key = ("p", _key(*[a for P in shared for a in P]))
        with self._meta:
            gid = self._gid.get(key)
            if gid is None:
                gid = self._gid[key] = len(self._data)
                self._data.append(self._marshal(kind, shared))
            self._ident[w][(kind, ident)] = gid
        return gid

    def _marshal(self, kind, shared):
        if not self._device_engine:
            return shared
        from .engine import GaviArrays, NodeArrays
        if kind == "verify":
            return NodeArrays(*shared)
        if kind in ("gavi", "comp"):
            return GaviArrays(shared)
        return shared

    # ---- the one thread that talks to the engine (calls on a handle are serialised, include/qpn_cuda.h) -----
    def _dispatch(self):
        while True:
            with self._cv:
                while not self._queue and not self._stop:
                    self._cv.wait()
                if self._stop and not self._queue:
                    return
                batch, self._queue = self._queue, []
            groups = {}
            for r in batch:
                groups.setdefault((r.kind, r.gid, tuple(sorted(r.kwargs.items()))), []).append(r)
            for (kind, gid, _), reqs in groups.items():
                try:
                    data = self._data[gid]
                    sizes = [len(r.stacked[0]) for r in reqs]
                    args = [np.vstack([r.stacked[k] for r in reqs]) if len(reqs) > 1 else reqs[0].stacked[k]
                            for k in range(len(reqs[0].stacked))]
                    kw = reqs[0].kwargs
                    self.device_calls += 1
                    if kind == "verify":
                        out = self.engine.verify_solution(data, args[0], **kw)
                        parts = _split(out, sizes)
                    elif kind == "gavi":
                        out = self.engine.gavi_solve(data, args[0], args[1], **kw)
                        keys = list(out)
                        parts = [dict(zip(keys, p)) for p in _split(tuple(out[k] for k in keys), sizes)]
                    elif kind == "comp":
                        parts = [p[0] for p in _split((self.engine.comp_indices(data, args[0], args[1], **kw),), sizes)]
                    else:
                        parts = [p[0] for p in _split((self.engine.halfspace_in(data, args[0], **kw),), sizes)]
                    for r, part in zip(reqs, parts):
                        r.result = part
                except Exception as e:                   # noqa: BLE001 -- delivered to every worker of the group
                    for r in reqs:
                        r.error = f"{type(e).__name__}: {e}"
                for r in reqs:
                    r.ready.set()

    def _execute(self, w, kind, ident, shared, stacked, kwargs):
        gid = self._register(w, kind, ident, shared) if shared is not None else self._ident[w][(kind, ident)]
        req = _Pending(kind, gid, stacked, kwargs)
        with self._cv:
            self.worker_calls += 1
            self._queue.append(req)
            self._cv.notify()
        req.ready.wait()
        if req.error is not None:
            raise RuntimeError(req.error)
        return req.result

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
                    conn.send(("err", str(e)))
        except (EOFError, OSError) as e:
            out[w] = RuntimeError(f"worker {w} died: {e}")

    def solve(self, X, keep_sol=False, stats=None):
        """solve(qpn, inits) for the whole batch; results in instance order.  `Sol` (the top level's solution
        pieces) stays in the workers unless keep_sol: it is large and a batch caller reads x_opt / solved."""
        if self.broken:
            raise RuntimeError("MultilevelPool: a worker failed in an earlier batch; create a new pool")
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        used = min(self.workers, len(X))
        out = [None] * used
        calls0, wcalls0 = self.device_calls, self.worker_calls
        threads = [threading.Thread(target=self._serve, args=(w, X[slice(*shard_range(len(X), w, used))], keep_sol, out), daemon=True)
                   for w in range(used)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for o in out:
            if isinstance(o, Exception):
                self.broken = True       # a worker that failed mid-batch has lost its place in the protocol
                raise o
        if stats is not None:
            for _, st in out:
                for k, v in st.items():
                    stats[k] = stats.get(k, 0) + v
            stats["workers"] = used
            stats["worker_calls"] = self.worker_calls - wcalls0
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
        with self._cv:
            self._stop = True
            self._cv.notify()
        self._dispatcher.join(timeout=10)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _split(arrays, sizes):
    """Row ranges of every array of a batched result, one tuple per request."""
    parts, o = [], 0
    for sz in sizes:
        parts.append(tuple(a[o:o + sz] if isinstance(a, np.ndarray) else a for a in arrays))
        o += sz
    return parts


def solve_multilevel_workers(qpn, X, workers, engine=None, device=0, chunk=256, stats=None, keep_sol=False):
    """One-shot form: a pool for this batch only."""
    with MultilevelPool(qpn, min(int(workers), len(X)), engine=engine, device=device, chunk=chunk) as pool:
        return pool.solve(X, keep_sol=keep_sol, stats=stats)
