"""Batched host state machine for networks with children.

`solve_base!` (/root/reference/src/algorithm.jl:1-127) is a data-dependent recursion: which node is
verified against which child pieces, which LP is asked, whether `solve_qep` runs -- all depends on
the instance.  Instead of rewriting that recursion in lockstep form, every instance of a batch runs
the unchanged per-instance code (`NetSolver`) on its own thread against a `BatchingEngine`: a proxy
with the engine's interface whose calls block until EVERY live instance is blocked on a call of its
own.  The scheduler then regroups the pending requests by (kind, shared problem data) -- instances
that verify the same node against the same pieces, solve the same level GAVI, classify against the
same node GAVI -- issues ONE batched device call per group, scatters the results and releases the
threads.  The device work of B instances thus costs one launch per distinct (kind, data) per round,
not one per instance; cycle detection, iterate caches and failure handling stay per instance
because they live in the per-instance `NetSolver` state.
"""
import threading
import time

import numpy as np


def _key(*arrays):
    return tuple((a.shape, a.tobytes()) if isinstance(a, np.ndarray) else a for a in arrays)


def _gavi_key(g):
    return _key(*[np.ascontiguousarray(g[k]) for k in ("M", "N", "o", "l1", "u1", "A", "B", "l2", "u2")])


class _Request:
    __slots__ = ("kind", "key", "shared", "args", "kwargs", "result", "error", "ready")

    def __init__(self, kind, key, shared, args, kwargs):
        self.kind, self.key, self.shared, self.args, self.kwargs = kind, key, shared, args, kwargs
        self.result, self.error = None, None
        self.ready = threading.Event()                   # each thread sleeps on its own event: no thundering herd


def plain_once(cache, key, compute):
    """Memoise compute() under key (single-threaded form)."""
    if key not in cache:
        cache[key] = compute()
    return cache[key]


class _Flight:
    __slots__ = ("value", "error", "done", "waiters", "event")

    def __init__(self):
        self.value, self.error, self.done, self.waiters, self.event = None, None, False, 0, threading.Event()


class BatchingEngine:
    """Engine proxy shared by the instance threads of one batch."""

    def __init__(self, engine, stall_seconds=120.0):
        self.engine = engine
        self.cv = threading.Condition()
        self.pending = []
        self.alive = 0
        self.stall_seconds = stall_seconds
        self.rounds = 0
        self.device_calls = 0
        self.requests = 0
        self._gavi_arrays = {}           # shared problem data marshalled once per distinct GAVI / node
        self._node_arrays = {}

    # ---- the engine interface, as the per-instance code calls it (single-instance arguments) -----------
    def verify_solution(self, node, x, tol=1e-4):
        node = tuple(np.ascontiguousarray(a) for a in node)
        return self._call("verify", _key(*node, tol), node, (np.atleast_2d(x),), dict(tol=tol))

    def gavi_solve(self, g, w, z0, presolve=True, max_pivots=0):
        return self._call("gavi", _key(_gavi_key(g), presolve, max_pivots), g, (np.atleast_2d(w), np.atleast_2d(z0)),
                          dict(presolve=presolve, max_pivots=max_pivots))

    def comp_indices(self, g, z, w, tol=1e-2):
        return self._call("comp", _key(_gavi_key(g), tol), g, (np.atleast_2d(z), np.atleast_2d(w)), dict(tol=tol))

    def halfspace_in(self, polys, x, tol=1e-6):
        polys = [tuple(np.ascontiguousarray(a) for a in P) for P in polys]
        return self._call("in", _key(*[a for P in polys for a in P], tol), polys, (np.atleast_2d(x),), dict(tol=tol))

    @property
    def launches(self):
        return self.engine.launches

    def once(self, cache, key, compute):
        """Memoise compute() under key across the instance threads: the first thread to ask computes (and may
        itself block on device calls), the others sleep until the value is there.  A sleeping thread is taken
        out of the rendezvous count, so the scheduler does not wait for a request it will not make; the owner
        puts its waiters back before it wakes them."""
        with self.cv:
            fl = cache.get(key)
            owner = fl is None
            if owner:
                fl = cache[key] = _Flight()
            elif not fl.done:
                fl.waiters += 1
                self.alive -= 1
                self.cv.notify()
        if owner:
            try:
                fl.value = compute()
            except BaseException as e:                   # noqa: BLE001
                fl.error = e
            with self.cv:
                fl.done = True
                self.alive += fl.waiters
            fl.event.set()
        elif not fl.done or not fl.event.is_set():
            fl.event.wait()
        if fl.error is not None:
            raise type(fl.error)(*fl.error.args)
        return fl.value

    # ---- rendezvous --------------------------------------------------------------------------------------
    def _call(self, kind, key, shared, args, kwargs):
        req = _Request(kind, key, shared, args, kwargs)
        with self.cv:
            self.pending.append(req)
            if len(self.pending) >= self.alive:
                self.cv.notify()                         # only the scheduler waits on cv
        req.ready.wait()
        if req.error is not None:
            raise type(req.error)(*req.error.args)          # a fresh object per thread (tracebacks do not pile up)
        return req.result

    def _execute(self, batch):
        """One device call per (kind, key) group; results scattered to the requests."""
        groups = {}
        for r in batch:
            groups.setdefault((r.kind, r.key), []).append(r)
        for (kind, key), reqs in groups.items():
            try:
                sizes = [len(r.args[0]) for r in reqs]
                stacked = [np.vstack([r.args[k] for r in reqs]) for k in range(len(reqs[0].args))]
                kw = reqs[0].kwargs
                self.device_calls += 1
                remote = getattr(self.engine, "remote_call", None)
                if remote is not None:
                    # a worker process (workers.py): the regrouped call goes to the process that owns the GPU
                    out = remote(kind, key, reqs[0].shared, stacked, kw)
                    if kind in ("verify",):
                        parts = self._split_tuple(out, sizes)
                    elif kind == "gavi":
                        parts = self._split_dict(out, sizes)
                    else:
                        parts = self._split_tuple((out,), sizes, single=True)
                elif kind == "verify":
                    na = self._node_arrays.get(key)
                    if na is None:
                        from .engine import Engine, NodeArrays
                        if len(self._node_arrays) > 4096:
                            self._node_arrays.clear()
                        na = self._node_arrays[key] = NodeArrays(*reqs[0].shared) if isinstance(self.engine, Engine) else reqs[0].shared
                    out = self.engine.verify_solution(na, stacked[0], **kw)
                    parts = self._split_tuple(out, sizes)
                elif kind == "gavi":
                    out = self.engine.gavi_solve(self._gavi(key, reqs[0].shared), stacked[0], stacked[1], **kw)
                    parts = self._split_dict(out, sizes)
                elif kind == "comp":
                    out = self.engine.comp_indices(self._gavi(key, reqs[0].shared), stacked[0], stacked[1], **kw)
                    parts = self._split_tuple((out,), sizes, single=True)
                else:
                    out = self.engine.halfspace_in(reqs[0].shared, stacked[0], **kw)
                    parts = self._split_tuple((out,), sizes, single=True)
                for r, p in zip(reqs, parts):
                    r.result = p
            except Exception as e:                       # noqa: BLE001 -- delivered to every thread of the group
                for r in reqs:
                    r.error = e

    def _gavi(self, key, g):
        ga = self._gavi_arrays.get(key[0])
        if ga is None:
            from .engine import Engine, GaviArrays
            if len(self._gavi_arrays) > 4096:
                self._gavi_arrays.clear()
            # marshalled (column-major copies) once per distinct GAVI when the device engine is behind this proxy
            ga = self._gavi_arrays[key[0]] = GaviArrays(g) if isinstance(self.engine, Engine) else g
        return ga

    @staticmethod
    def _split_tuple(out, sizes, single=False):
        parts, o = [], 0
        for s in sizes:
            piece = tuple(a[o:o + s] for a in out)
            parts.append(piece[0] if single else piece)
            o += s
        return parts

    @staticmethod
    def _split_dict(out, sizes):
        parts, o = [], 0
        for s in sizes:
            parts.append({k: (v[o:o + s] if isinstance(v, np.ndarray) else v) for k, v in out.items()})
            o += s
        return parts

    def run(self, jobs):
        """jobs: callables, one per instance.  Returns their results in order (an exception raised by a job
        is re-raised here after every thread has ended)."""
        results, errors = [None] * len(jobs), [None] * len(jobs)

        def worker(i):
            try:
                results[i] = jobs[i]()
            except BaseException as e:                   # noqa: BLE001
                errors[i] = e
            finally:
                with self.cv:
                    self.alive -= 1
                    self.cv.notify()

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(len(jobs))]
        with self.cv:
            self.alive = len(jobs)
        for t in threads:
            t.start()
        last_progress = time.monotonic()
        while True:
            with self.cv:
                while self.alive > 0 and len(self.pending) < self.alive:
                    if not self.cv.wait(timeout=1.0) and time.monotonic() - last_progress > self.stall_seconds:
                        raise RuntimeError("BatchingEngine: no progress (an instance thread is stuck outside the engine)")
                if self.alive == 0 and not self.pending:
                    break
                batch, self.pending = self.pending, []
            self.rounds += 1
            self.requests += len(batch)
            self._execute(batch)                          # device work outside the lock
            last_progress = time.monotonic()
            for r in batch:
                r.ready.set()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        return results
