"""`solve(qpn, inits::Matrix)` for networks with children through the native state machine of libqpn_cuda
(`qpn_net_*`, include/qpn_cuda.h; sources in csrc/net/).

The QPNet is flattened to the plain arrays of `qpn_net_desc` -- the same arrays a Julia `QPNet` holds after
`setup(:name)` (/root/reference/src/programs.jl:79-116) -- and handed over once; a batch then costs one call.
`NetBinding` is written against a symbol prefix so that the test suite can drive its checker build of the same state
machine (the host logic of csrc/net/ over CPU numerics, built under the test infrastructure) through the identical
marshalling code; the product only ever passes "qpn_net_".
"""
import ctypes as C

import numpy as np

from .algorithm import projection_vectors

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
ubp = C.POINTER(C.c_uint8)

ERRORS = {0: "", 1: "Cycling detected (noticed solution iterate returned to a previous value).",
          2: "AVI solve error. This might be because one of the qps is unbounded or ill-conditioned.",
          3: "Detected disagreement in solution status between qp solution processer and equilibrium solver.",
          4: "Can't find solution", 5: "This shouldn't happen. Solution graph is empty.",
          6: "comp_indices: an index belongs to no set (the reference's @assert)", 7: "Too many solutions to combine.",
          8: "Solution graphs were not properly populated.", 9: "Cycling check requested, but num_projections == 0."}


class QpnNetDesc(C.Structure):
    _fields_ = [("nv", C.c_int32), ("nplayers", C.c_int32), ("nlevels", C.c_int32), ("npolys", C.c_int32),
                ("Q", dp), ("q", dp), ("var_ptr", ip), ("var_idx", ip), ("con_ptr", ip), ("con_idx", ip),
                ("child_ptr", ip), ("child_idx", ip), ("level_of", ip), ("poly_ptr", ip),
                ("poly_A", dp), ("poly_l", dp), ("poly_u", dp),
                ("max_iters", C.c_int32), ("num_projections", C.c_int32), ("exploration_vertices", C.c_int32),
                ("gen_solution_map", C.c_int32), ("check_for_cycling", C.c_int32),
                ("remove_subsets_at", ubp), ("proj", dp)]


def _csr(lists):
    ptr = np.zeros(len(lists) + 1, np.int32)
    for k, l in enumerate(lists):
        ptr[k + 1] = ptr[k] + len(l)
    idx = np.array([v for l in lists for v in l], np.int32) if ptr[-1] else np.zeros(1, np.int32)
    return ptr, idx


def build_net_desc(qpn):
    """QPNet -> (qpn_net_desc, arrays kept alive).  Players and constraint polys become 0-based, in id order."""
    pids = sorted(qpn.qps)
    cids = sorted(qpn.constraints)
    ppos = {p: k for k, p in enumerate(pids)}
    cpos = {c: k for k, c in enumerate(cids)}
    nv, npl = qpn.n_vars, len(pids)
    nl = qpn.num_levels()
    keep = {}
    keep["Q"] = np.ascontiguousarray(np.stack([qpn.qps[p].Q for p in pids]), dtype=np.float64)
    keep["q"] = np.ascontiguousarray(np.stack([qpn.qps[p].q for p in pids]), dtype=np.float64)
    keep["var_ptr"], keep["var_idx"] = _csr([qpn.qps[p].var_indices for p in pids])
    keep["con_ptr"], keep["con_idx"] = _csr([[cpos[c] for c in qpn.qps[p].constraint_indices] for p in pids])
    keep["child_ptr"], keep["child_idx"] = _csr([[ppos[j] for j in sorted(qpn.network_edges[p])] for p in pids])
    level_of = np.zeros(npl, np.int32)
    for lv, players in qpn.network_depth_map.items():
        for p in players:
            level_of[ppos[p]] = lv - 1
    keep["level_of"] = level_of
    polys = [qpn.constraints[c] for c in cids]
    keep["poly_ptr"] = np.concatenate([[0], np.cumsum([len(P) for P in polys])]).astype(np.int32)
    rows = int(keep["poly_ptr"][-1])
    keep["poly_A"] = np.ascontiguousarray(np.vstack([P.A for P in polys]).reshape(rows, nv) if rows else np.zeros((1, nv)))
    keep["poly_l"] = np.ascontiguousarray(np.concatenate([P.l for P in polys]) if rows else np.zeros(1))
    keep["poly_u"] = np.ascontiguousarray(np.concatenate([P.u for P in polys]) if rows else np.zeros(1))
    opt = qpn.options
    lv = opt.levels_to_remove_subsets
    keep["remove"] = np.array([1 if (lv is None or (k + 1) in lv) else 0 for k in range(nl)], np.uint8)
    proj = projection_vectors(qpn)
    keep["proj"] = np.ascontiguousarray(proj, dtype=np.float64) if len(proj) else np.zeros((1, nv))
    p = lambda a, t=dp: a.ctypes.data_as(t)
    desc = QpnNetDesc(nv, npl, nl, len(polys), p(keep["Q"]), p(keep["q"]), p(keep["var_ptr"], ip), p(keep["var_idx"], ip),
                      p(keep["con_ptr"], ip), p(keep["con_idx"], ip), p(keep["child_ptr"], ip), p(keep["child_idx"], ip),
                      p(keep["level_of"], ip), p(keep["poly_ptr"], ip), p(keep["poly_A"]), p(keep["poly_l"]), p(keep["poly_u"]),
                      int(opt.max_iters), int(len(proj)), int(opt.exploration_vertices), int(bool(opt.gen_solution_map)),
                      int(bool(opt.check_for_cycling)), p(keep["remove"], ubp), p(keep["proj"]))
    return desc, keep, pids


class NetBinding:
    """One native net object.  lib: a loaded shared library; prefix: "qpn_net_" (CUDA engine; `handle` = qpn_handle*);
    a library without a handle argument (the tests' checker build) passes handle=None and its own prefix."""

    def __init__(self, qpn, lib, prefix="qpn_net_", handle=None, threads=None):
        from .model import Poly
        self._Poly = Poly
        self.lib, self.prefix, self.qpn = lib, prefix, qpn
        self.desc, self._keep, self.pids = build_net_desc(qpn)
        self.nv, self.nlevels = qpn.n_vars, qpn.num_levels()
        ptr = C.c_void_p()
        create = self._f("create")
        rc = create(handle, C.byref(self.desc), C.byref(ptr)) if handle is not None else create(C.byref(self.desc), C.byref(ptr))
        if rc != 0:
            raise RuntimeError(f"{prefix}create failed")
        self.ptr = ptr
        if threads:
            self.set_option("threads", threads)

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def set_option(self, name, value):
        if self._f("set_option")(self.ptr, name.encode(), C.c_int64(int(value))) != 0:
            raise ValueError(f"unknown net option {name}")

    def close(self):
        if getattr(self, "ptr", None):
            self._f("destroy")(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve_arrays(self, inits, out=None):
        """inits (B, nv) -> dict of arrays: x (B, nv), solved (B,), level_iters (B, nlevels), error (B,).
        out: preallocated arrays of those names (e.g. pinned memory)."""
        X = inits if (isinstance(inits, np.ndarray) and inits.ndim == 2 and inits.dtype == np.float64 and inits.flags.c_contiguous) \
            else np.ascontiguousarray(np.atleast_2d(inits), dtype=np.float64)
        B = X.shape[0]
        if out is not None:
            x, solved, iters, err = out["x"], out["solved"], out["level_iters"], out["error"]
        else:
            x = np.empty((B, self.nv)); solved = np.zeros(B, np.uint8)
            iters = np.zeros((B, self.nlevels), np.int32); err = np.zeros(B, np.int32)
        rc = self._f("solve_batched")(self.ptr, B, X.ctypes.data_as(dp), x.ctypes.data_as(dp), solved.ctypes.data_as(ubp),
                                      iters.ctypes.data_as(ip), err.ctypes.data_as(ip))
        if rc != 0:
            msg = ""
            if self.prefix == "qpn_net_":
                self.lib.qpn_net_last_error.restype = C.c_char_p
                msg = (self.lib.qpn_net_last_error(self.ptr) or b"").decode()
            from .engine import EngineError
            raise EngineError(f"{self.prefix}solve_batched failed: {msg}")
        return dict(x=x, solved=solved.view(bool) if solved.dtype == np.uint8 else solved, level_iters=iters, error=err)

    def solve_dev(self, batch, inits_ptr, x_out_ptr, out=None):
        """Device-resident form (qpn_net_solve_batched_dev): inits / x_out are device pointers (nv x batch)."""
        B = int(batch)
        out = out or dict(solved=np.zeros(B, np.uint8), level_iters=np.zeros((B, self.nlevels), np.int32), error=np.zeros(B, np.int32))
        rc = self._f("solve_batched_dev")(self.ptr, B, C.c_void_p(int(inits_ptr)), C.c_void_p(int(x_out_ptr)), out["solved"].ctypes.data_as(ubp),
                                          out["level_iters"].ctypes.data_as(ip), out["error"].ctypes.data_as(ip))
        if rc != 0:
            from .engine import EngineError
            self.lib.qpn_net_last_error.restype = C.c_char_p
            raise EngineError("qpn_net_solve_batched_dev failed: " + (self.lib.qpn_net_last_error(self.ptr) or b"").decode())
        return out

    def profile(self):
        """qpn_net_profile: launches / units / summed ms per kernel kind, staged bytes per direction."""
        o = np.zeros(24)
        self._f("profile")(self.ptr, o.ctypes.data_as(dp))
        kinds = ("verify", "solve_qep", "member", "group", "cycle")
        d = {k: dict(launches=int(o[4 * i]), units=int(o[4 * i + 1]), ms=float(o[4 * i + 2])) for i, k in enumerate(kinds)}
        d["h2d_bytes"], d["d2h_bytes"] = int(o[20]), int(o[21])
        return d

    def piece(self, pid):
        m = self._f("piece_rows")(self.ptr, int(pid))
        A = np.zeros((max(m, 1), self.nv)); l = np.zeros(max(m, 1)); u = np.zeros(max(m, 1))
        rl = np.zeros(max(m, 1), np.uint8); ru = np.zeros(max(m, 1), np.uint8)
        self._f("piece_get")(self.ptr, int(pid), A.ctypes.data_as(dp), l.ctypes.data_as(dp), u.ctypes.data_as(dp),
                             rl.ctypes.data_as(ubp), ru.ctypes.data_as(ubp))
        return self._Poly(A[:m], l[:m], u[:m], rl[:m].astype(bool), ru[:m].astype(bool), normalize=False)

    def sol(self, b):
        """ret.Sol of instance b of the last batch: {player id: [Poly, ...]} (algorithm.jl:116)."""
        out = {}
        for k, p in enumerate(self.pids):
            n = self._f("sol_count")(self.ptr, int(b), k)
            out[p] = [self.piece(self._f("sol_piece")(self.ptr, int(b), k, j)) for j in range(n)] if n >= 0 else None
        return out

    def stats(self):
        out = np.zeros(16, np.int64)
        self._f("stats")(self.ptr, out.ctypes.data_as(C.POINTER(C.c_int64)))
        names = ["launches", "rounds", "requests", "calls", "lps", "pieces", "nodes", "gavis", "collect_misses", "combine_misses",
                 "host_ns", "backend_ns", "cohort_splits", "apply_ns", "lps_empty", "lp_calls"]
        return {n: int(v) for n, v in zip(names, out)}

    def solve(self, inits, keep_sol=False):
        """List of result dicts with the reference's fields (algorithm.jl:116,125)."""
        ret = self.solve_arrays(inits)
        outs = []
        for b in range(len(ret["solved"])):
            if ret["solved"][b]:
                r = dict(solved=True, x_opt=ret["x"][b].copy(), identified_request=set(), x_alts=[], level_iters=ret["level_iters"][b].tolist())
                if keep_sol:
                    r["Sol"] = self.sol(b)
                outs.append(r)
            else:
                outs.append(dict(solved=False, x_fail=ret["x"][b].copy(), x_opt=None, error=ERRORS.get(int(ret["error"][b]) & 0xff, "?"), error_code=int(ret["error"][b]),
                                 level_iters=ret["level_iters"][b].tolist()))
        return outs
