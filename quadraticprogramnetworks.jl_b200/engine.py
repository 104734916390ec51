"""ctypes binding of libqpn_cuda (include/qpn_cuda.h) -- the same symbols the Julia
wrapper reaches with `ccall`.  There is no CPU fallback: if the shared
library or a B200 is missing this module raises.

Array conventions on this side: batched vectors are numpy arrays of shape (batch, n),
C-contiguous -- byte-for-byte the `n x batch` column-major layout the C ABI takes.
Matrices are given in math layout M[i, j] and converted to column-major here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPN_CUDA_LIB") or os.path.join(_HERE, "lib", "libqpn_cuda.so")   # override: debug builds only

SUCCESS, RAY_TERM, MAX_ITERS, FAILURE = 1, 2, 3, 4

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
bp = C.POINTER(C.c_int8)
ubp = C.POINTER(C.c_uint8)


class QpnMatrix(C.Structure):
    _fields_ = [("dense", dp), ("colptr", ip), ("rowval", ip), ("nzval", dp),
                ("nnz", C.c_int32), ("index_base", C.c_int32), ("is_shared", C.c_int32)]


class QpnGavi(C.Structure):
    _fields_ = [("d1", C.c_int32), ("d2", C.c_int32), ("np", C.c_int32),
                ("M", dp), ("N", dp), ("o", dp), ("l1", dp), ("u1", dp),
                ("A", dp), ("B", dp), ("l2", dp), ("u2", dp)]


class QpnNode(C.Structure):
    _fields_ = [("nd", C.c_int32), ("nv", C.c_int32), ("m", C.c_int32),
                ("Qd", dp), ("qd", dp), ("A", dp), ("l", dp), ("u", dp), ("dec", ip)]


class QpnLevel(C.Structure):
    _fields_ = [("nv", C.c_int32), ("nplayers", C.c_int32), ("players", C.POINTER(QpnNode)),
                ("gavi", QpnGavi), ("dec", ip), ("nd_level", C.c_int32), ("par", ip),
                ("max_iters", C.c_int32), ("num_projections", C.c_int32), ("proj", dp)]


class EngineError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libqpn_cuda.so or fail loudly (the product path has no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.qpn_last_error.restype = C.c_char_p
    lib.qpn_last_error.argtypes = [C.c_void_p]
    lib.qpn_launch_count.restype = C.c_int64
    lib.qpn_launch_count.argtypes = [C.c_void_p]
    lib.qpn_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.qpn_big_launch_count.restype = C.c_int64
    lib.qpn_big_launch_count.argtypes = [C.c_void_p]
    for name in EXPORTS:
        getattr(lib, name)   # every declared symbol must resolve
    # the ctypes mirrors of the ABI structs must have the library's layout
    sizes = (C.c_int32 * 6)()
    lib.qpn_abi_struct_sizes(sizes)
    from .netsolve import QpnNetDesc
    mine = [C.sizeof(QpnMatrix), C.sizeof(QpnGavi), C.sizeof(QpnNode), C.sizeof(QpnLevel), C.sizeof(QpnNetDesc)]
    if list(sizes)[:5] != mine:
        raise EngineError(f"ABI struct sizes differ: library {list(sizes)[:5]}, ctypes mirrors {mine}")
    _lib = lib
    return lib


# Every symbol include/qpn_cuda.h declares (tests check the .so exports them all).
EXPORTS = [
    "qpn_create", "qpn_destroy", "qpn_last_error", "qpn_device", "qpn_launch_count", "qpn_synchronize",
    "qpn_avi_solve_batched", "qpn_avi_solve_batched_dev", "qpn_check_avi_batched",
    "qpn_gavi_solve_batched", "qpn_gavi_solve_batched_dev", "qpn_comp_indices_batched",
    "qpn_halfspace_in_batched", "qpn_verify_solution_batched",
    "qpn_level_equilibrium_batched", "qpn_level_equilibrium_batched_dev",
    "qpn_level_upload", "qpn_level_release", "qpn_level_equilibrium_resident", "qpn_level_equilibrium_resident_dev",
    "qpn_malloc", "qpn_free", "qpn_memcpy_h2d", "qpn_memcpy_d2h", "qpn_set_option", "qpn_big_launch_count", "qpn_level_info",
    "qpn_net_create", "qpn_net_destroy", "qpn_net_last_error", "qpn_net_set_option", "qpn_net_solve_batched",
    "qpn_net_solve_batched_dev", "qpn_net_profile", "qpn_host_register", "qpn_host_unregister", "qpn_abi_struct_sizes",
    "qpn_net_sol_count", "qpn_net_sol_piece", "qpn_net_piece_rows", "qpn_net_piece_get", "qpn_net_stats",
]


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def _colmajor(M):
    """math-layout matrix (..., r, c) -> bytes of column-major storage."""
    M = np.asarray(M, dtype=np.float64)
    return np.ascontiguousarray(np.swapaxes(M, -1, -2))


def _p(a, t=dp):
    return None if a is None else a.ctypes.data_as(t)


# `ndarray.ctypes.data_as` costs ~5 us per call -- six of them were a quarter of a 4,096-instance level call.
# Buffers that come back call after call (pinned batches, preallocated outputs) are looked up by identity instead;
# the cache keeps the array alive, so an id cannot be recycled while its entry exists.
_PTR_CACHE = {}
_PTR_CACHE_MAX = 16


def _vp(a):
    """c_void_p of a C-contiguous numpy array (None passes through), memoised per array object."""
    if a is None:
        return None
    hit = _PTR_CACHE.get(id(a))
    if hit is not None and hit[0] is a:
        return hit[1]
    if len(_PTR_CACHE) >= _PTR_CACHE_MAX:
        _PTR_CACHE.clear()
    ptr = C.c_void_p(a.__array_interface__["data"][0])
    _PTR_CACHE[id(a)] = (a, ptr)
    return ptr


class GaviArrays:
    """Keeps the column-major copies of a GAVI's blocks alive next to the C struct."""

    def __init__(self, g):
        self.d1, self.d2, self.np_ = len(g["l1"]), len(g["l2"]), g["N"].shape[1]
        self.keep = {k: (_colmajor(g[k]) if g[k].ndim == 2 else _c(g[k])) for k in ("M", "N", "o", "l1", "u1", "A", "B", "l2", "u2")}
        k = self.keep
        self.struct = QpnGavi(self.d1, self.d2, self.np_, _p(k["M"]), _p(k["N"]), _p(k["o"]), _p(k["l1"]), _p(k["u1"]),
                              _p(k["A"]), _p(k["B"]), _p(k["l2"]), _p(k["u2"]))


class NodeArrays:
    def __init__(self, Qd, qd, A, l, u, dec):
        Qd = np.atleast_2d(np.asarray(Qd, dtype=np.float64))
        self.nd, self.nv = Qd.shape
        A = np.asarray(A, dtype=np.float64).reshape(-1, self.nv)
        self.m = A.shape[0]
        self.keep = dict(Qd=_colmajor(Qd), qd=_c(qd), A=_colmajor(A), l=_c(l), u=_c(u), dec=_c(dec, np.int32))
        k = self.keep
        self.struct = QpnNode(self.nd, self.nv, self.m, _p(k["Qd"]), _p(k["qd"]), _p(k["A"]), _p(k["l"]), _p(k["u"]), _p(k["dec"], ip))


class Engine:
    """One handle per GPU (include/qpn_cuda.h: calls on a handle are serialised by the caller)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.qpn_create(int(device), C.byref(h))
        if rc != 0:
            raise EngineError(self.lib.qpn_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.qpn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.qpn_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.qpn_launch_count(self.h))

    @property
    def big_launches(self):
        """Launches that took the global-memory tableau path."""
        return int(self.lib.qpn_big_launch_count(self.h))

    def synchronize(self):
        self._ck(self.lib.qpn_synchronize(self.h))

    def host_register(self, a):
        """Page-lock a numpy array the caller owns (qpn_host_register): the host-pointer entry points then read / write
        it in place.  Returns the array; call host_unregister before it is freed."""
        self._ck(self.lib.qpn_host_register(self.h, C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))
        return a

    def host_unregister(self, a):
        self._ck(self.lib.qpn_host_unregister(self.h, C.c_void_p(a.ctypes.data)))

    def set_option(self, name, value):
        """qpn_set_option: "force_big" (global-memory tableau path for every solve), "big_ctas_per_sm",
        "big_slot_in_smem" (0: never keep a compact tableau slot in shared memory)."""
        self._ck(self.lib.qpn_set_option(self.h, name.encode(), C.c_int64(int(value))))

    # ---- matrices ----------------------------------------------------------------------
    @staticmethod
    def matrix(M=None, csc=None, index_base=0):
        """dense: M (n,n) shared or (B,n,n).  csc: (colptr, rowval, nzval) with nzval (nnz,) or (B,nnz)."""
        keep = {}
        if M is not None:
            M = np.asarray(M, dtype=np.float64)
            keep["dense"] = _colmajor(M)
            st = QpnMatrix(_p(keep["dense"]), None, None, None, 0, 0, int(M.ndim == 2))
        else:
            colptr, rowval, nzval = csc
            keep["colptr"], keep["rowval"], keep["nzval"] = _c(colptr, np.int32), _c(rowval, np.int32), _c(nzval)
            st = QpnMatrix(None, _p(keep["colptr"], ip), _p(keep["rowval"], ip), _p(keep["nzval"]),
                           len(keep["rowval"]), int(index_base), int(keep["nzval"].ndim == 1))
        return st, keep

    # ---- solve_avi (avi.jl:63-77) ------------------------------------------------------
    def avi_solve(self, M, q, l, u, z0, max_pivots=0, csc=None, index_base=0, want_basis=True):
        q, z0 = _c(q), _c(z0)
        B, n = q.shape
        st, keep = self.matrix(M, csc, index_base)
        l, u = _c(l), _c(u)
        z = np.empty((B, n))
        status = np.empty(B, np.int32)
        piv = np.empty(B, np.int32)
        basis = np.empty((B, n), np.int8) if want_basis else None
        self._ck(self.lib.qpn_avi_solve_batched(self.h, n, B, C.byref(st), _p(q), _p(l), _p(u), int(l.ndim == 1), _p(z0),
                                                int(max_pivots), _p(z), _p(status, ip), _p(piv, ip), _p(basis, bp)))
        return z, status, piv, basis

    # ---- check_avi_solution (avi.jl:148-156) -------------------------------------------
    def check_avi(self, M, q, l, u, z, tol=1e-6, csc=None, index_base=0):
        q, z = _c(q), _c(z)
        B, n = q.shape
        st, keep = self.matrix(M, csc, index_base)
        l, u = _c(l), _c(u)
        bad = np.empty(B, np.int32)
        r = np.empty((B, n))
        self._ck(self.lib.qpn_check_avi_batched(self.h, n, B, C.byref(st), _p(q), _p(l), _p(u), int(l.ndim == 1), _p(z),
                                                C.c_double(tol), _p(bad, ip), _p(r)))
        return bad, r

    # ---- solve_gavi (avi.jl:101-111) ---------------------------------------------------
    def gavi_solve(self, g, w, z0, presolve=True, max_pivots=0):
        ga = g if isinstance(g, GaviArrays) else GaviArrays(g)
        w = _c(w).reshape(-1, ga.np_) if ga.np_ else np.zeros((len(z0), 0))
        z0 = _c(z0)
        B = z0.shape[0]
        dz, n = ga.d1 + ga.d2, ga.d1 + 2 * ga.d2
        z = np.empty((B, dz)); zf = np.empty((B, n))
        status = np.empty(B, np.int32); piv = np.empty(B, np.int32); basis = np.empty((B, n), np.int8)
        self._ck(self.lib.qpn_gavi_solve_batched(self.h, C.byref(ga.struct), B, _p(w), _p(z0), int(presolve), int(max_pivots),
                                                 _p(z), _p(zf), _p(status, ip), _p(piv, ip), _p(basis, bp)))
        return dict(z=z, z_full=zf, status=status, pivots=piv, basis=basis)

    # ---- comp_indices (avi_solutions.jl:511-612) ---------------------------------------
    def comp_indices(self, g, z, w, tol=1e-2):
        ga = g if isinstance(g, GaviArrays) else GaviArrays(g)
        z = _c(z)
        B = z.shape[0]
        w = _c(w).reshape(B, ga.np_)
        mask = np.empty((B, ga.d1 + ga.d2), np.int8)
        self._ck(self.lib.qpn_comp_indices_batched(self.h, C.byref(ga.struct), B, _p(z), _p(w), C.c_double(tol), _p(mask, bp)))
        return mask

    # ---- Base.in (sets.jl:820-853) -----------------------------------------------------
    def halfspace_in(self, polys, x, tol=1e-6):
        """polys: list of (A, l, u[, rl, ru]) over the same dimension; x: (npts, d).
        Returns bool array (npts, npoly)."""
        x = _c(x)
        npts, d = x.shape
        ptr = np.zeros(len(polys) + 1, np.int32)
        for k, P in enumerate(polys):
            ptr[k + 1] = ptr[k] + len(P[1])
        A = np.vstack([np.asarray(P[0], dtype=np.float64).reshape(-1, d) for P in polys]) if ptr[-1] else np.zeros((0, d))
        l = np.concatenate([np.asarray(P[1], dtype=np.float64) for P in polys]) if ptr[-1] else np.zeros(0)
        u = np.concatenate([np.asarray(P[2], dtype=np.float64) for P in polys]) if ptr[-1] else np.zeros(0)
        rl = np.concatenate([np.asarray(P[3] if len(P) > 3 else np.zeros(len(P[1])), dtype=np.uint8) for P in polys]) if ptr[-1] else np.zeros(0, np.uint8)
        ru = np.concatenate([np.asarray(P[4] if len(P) > 4 else np.zeros(len(P[1])), dtype=np.uint8) for P in polys]) if ptr[-1] else np.zeros(0, np.uint8)
        Ac = _colmajor(A)
        out = np.empty((npts, len(polys)), np.uint8)
        self._ck(self.lib.qpn_halfspace_in_batched(self.h, len(polys), d, int(ptr[-1]), _p(ptr, ip), _p(Ac), _p(_c(l)), _p(_c(u)),
                                                   _p(_c(rl, np.uint8), ubp), _p(_c(ru, np.uint8), ubp), npts, _p(x), C.c_double(tol), _p(out, ubp)))
        return out.astype(bool)

    # ---- verify_solution (qp_processing.jl:57-149) -------------------------------------
    def verify_solution(self, node, x, tol=1e-4):
        na = node if isinstance(node, NodeArrays) else NodeArrays(*node)
        x = _c(x)
        B = x.shape[0]
        sol = np.empty(B, np.uint8); lam = np.empty((B, na.m)); how = np.empty(B, np.int32); act = np.empty((B, na.m), np.int8)
        self._ck(self.lib.qpn_verify_solution_batched(self.h, C.byref(na.struct), B, _p(x), C.c_double(tol), _p(sol, ubp), _p(lam),
                                                      _p(how, ip), _p(act, bp)))
        return sol.astype(bool), lam, how, act

    # ---- fused level loop (algorithm.jl:13-118) ----------------------------------------
    def level_equilibrium(self, level, x_init, want_lam=True):
        """level: LevelArrays.  x_init: (B, nv)."""
        x_init = _c(x_init)
        B, nv = x_init.shape
        x = np.empty((B, nv)); solved = np.empty(B, np.uint8); iters = np.empty(B, np.int32); piv = np.empty(B, np.int32)
        lam = np.empty((B, level.lam_total)) if want_lam else None
        self._ck(self.lib.qpn_level_equilibrium_batched(self.h, C.byref(level.struct), B, _p(x_init), _p(x), _p(solved, ubp),
                                                        _p(iters, ip), _p(piv, ip), _p(lam)))
        return dict(x=x, solved=solved.astype(bool), iters=iters, pivots=piv, lam=lam)


class ResidentLevel:
    """A level whose problem data lives on the GPU (qpn_level_upload).  Batches then move only
    x_init in and the results out -- or nothing at all with the `_dev` form."""

    def __init__(self, engine, level):
        self.engine, self.level = engine, level
        self.lam_total, self.nv = level.lam_total, level.struct.nv
        ptr = C.c_void_p()
        engine._ck(engine.lib.qpn_level_upload(engine.h, C.byref(level.struct), C.byref(ptr)))
        self.ptr = ptr

    def info(self):
        """qpn_level_info: plan shapes and which tableau path the level runs on."""
        out = np.zeros(8, np.int32)
        self.engine._ck(self.engine.lib.qpn_level_info(self.engine.h, self.ptr, _p(out, ip)))
        return dict(n=int(out[0]), ncol0=int(out[1]), plan_pivots=int(out[2]), presolve_n=int(out[3]), presolve_ncol0=int(out[4]),
                    presolve_plan_pivots=int(out[5]), big=bool(out[6]))

    def release(self):
        if getattr(self, "ptr", None):
            self.engine.lib.qpn_level_release(self.engine.h, self.ptr)
            self.ptr = None

    def solve(self, x_init, out=None, want_lam=True):
        """Host buffers (numpy, ideally pinned): copies x_init in, results out, synchronises."""
        e = self.engine
        if not (isinstance(x_init, np.ndarray) and x_init.dtype == np.float64 and x_init.flags.c_contiguous):
            x_init = _c(x_init)
        B = x_init.shape[0]
        if out is None:
            out = dict(x=np.empty((B, self.nv)), solved=np.empty(B, np.uint8), iters=np.empty(B, np.int32),
                       pivots=np.empty(B, np.int32), lam=np.empty((B, self.lam_total)) if want_lam else None)
        rc = e.lib.qpn_level_equilibrium_resident(e.h, self.ptr, B, _vp(x_init), _vp(out["x"]), _vp(out["solved"]),
                                                  _vp(out["iters"]), _vp(out["pivots"]), _vp(out.get("lam")))
        if rc != 0:
            e._ck(rc)
        return out

    def solve_dev(self, batch, x_init_ptr, x_out_ptr, solved_ptr, iters_ptr, pivots_ptr, lam_ptr=None, stream=0):
        """Device pointers (e.g. torch tensors' data_ptr()); asynchronous on `stream`."""
        e = self.engine
        vp = lambda p: C.c_void_p(int(p)) if p else None
        e._ck(e.lib.qpn_level_equilibrium_resident_dev(e.h, self.ptr, int(batch), vp(x_init_ptr), vp(x_out_ptr), vp(solved_ptr),
                                                       vp(iters_ptr), vp(pivots_ptr), vp(lam_ptr), vp(stream)))


class LevelArrays:
    """A level without child pieces: players (NodeArrays), its GAVI, index maps, options."""

    def __init__(self, nv, nodes, gavi, dec, par, max_iters=150, proj=None):
        self.nodes = [n if isinstance(n, NodeArrays) else NodeArrays(*n) for n in nodes]
        self.gavi = gavi if isinstance(gavi, GaviArrays) else GaviArrays(gavi)
        self.dec = _c(dec, np.int32)
        self.par = _c(par, np.int32)
        self.proj = None if proj is None or len(proj) == 0 else _c(proj)     # (nproj, nv) rows = vectors
        self.node_array = (QpnNode * len(self.nodes))(*[n.struct for n in self.nodes])
        self.lam_total = sum(n.m for n in self.nodes)
        nproj = 0 if self.proj is None else self.proj.shape[0]
        self.struct = QpnLevel(nv, len(self.nodes), C.cast(self.node_array, C.POINTER(QpnNode)), self.gavi.struct,
                               _p(self.dec, ip), len(self.dec), _p(self.par, ip), int(max_iters), nproj, _p(self.proj))
