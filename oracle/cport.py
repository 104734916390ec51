"""ORACLE (test infrastructure): ctypes binding of oracle/_build/libqpn_oracle.so.

Build with `make -C oracle` (or __graft_entry__.build()).  Matrices are passed
column-major (np.asfortranarray), exactly as the C-ABI of libqpn_cuda takes them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqpn_oracle.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
bp = C.POINTER(C.c_int8)
ubp = C.POINTER(C.c_uint8)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "qpn_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.qpo_avi_solve.restype = C.c_int
        _lib.qpo_avi_solve_batched.restype = C.c_int
        _lib.qpo_gavi_solve.restype = C.c_int
        _lib.qpo_check_avi.restype = C.c_int
        _lib.qpo_halfspace_in.restype = C.c_int
        _lib.qpo_verify_solution.restype = C.c_int
    return _lib


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a, t=dp):
    return a.ctypes.data_as(t)


def avi_solve(M, q, l, u, z0, max_pivots=0):
    L = lib()
    M, q, l, u, z0 = _f(M), _f(q), _f(l), _f(u), _f(z0)
    n = len(q)
    z = np.zeros(n)
    st = np.zeros(1, np.int32)
    pv = np.zeros(1, np.int32)
    basis = np.zeros(n, np.int8)
    rc = L.qpo_avi_solve(n, _p(M), _p(q), _p(l), _p(u), _p(z0), int(max_pivots), _p(z), _p(st, ip), _p(pv, ip), _p(basis, bp))
    assert rc == 0
    return z, int(st[0]), int(pv[0]), basis


def avi_solve_batched(M, q, l, u, z0, max_pivots=0, threads=1):
    """M: (n,n) shared or (B,n,n) per instance (each n x n, math layout M[i,j]);
    q, z0: (B,n); l,u: (n,) shared or (B,n)."""
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.float64)
    z0 = np.ascontiguousarray(z0, dtype=np.float64)
    B, n = q.shape
    M = np.asarray(M, dtype=np.float64)
    shared = M.ndim == 2
    Mc = np.ascontiguousarray(M.T if shared else M.transpose(0, 2, 1))   # column-major per instance
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    lu_shared = l.ndim == 1
    z = np.zeros((B, n))
    st = np.zeros(B, np.int32)
    pv = np.zeros(B, np.int32)
    basis = np.zeros((B, n), np.int8)
    rc = L.qpo_avi_solve_batched(n, B, _p(Mc), int(shared), _p(q), _p(l), _p(u), int(lu_shared), _p(z0),
                                 int(max_pivots), _p(z), _p(st, ip), _p(pv, ip), _p(basis, bp), int(threads))
    assert rc == 0
    return z, st, pv, basis


def check_avi(M, q, l, u, z, tol=1e-6):
    L = lib()
    M, q, l, u, z = _f(M), _f(q), _f(l), _f(u), _f(z)
    r = np.zeros(len(q))
    bad = L.qpo_check_avi(len(q), _p(M), _p(q), _p(l), _p(u), _p(z), C.c_double(tol), _p(r))
    return bad > 0, bad, r


def gavi_solve(g, z0, w, presolve=True, max_pivots=0):
    L = lib()
    d1, d2 = len(g["l1"]), len(g["l2"])
    npar = g["N"].shape[1]
    arrs = [_f(g[k]) for k in ("M", "N", "o", "l1", "u1", "A", "B", "l2", "u2")]
    w = _f(w)
    z0 = _f(z0).copy()
    z = np.zeros(d1 + d2)
    zfull = np.zeros(d1 + 2 * d2)
    st = np.zeros(1, np.int32)
    pv = np.zeros(1, np.int32)
    basis = np.zeros(d1 + 2 * d2, np.int8)
    rc = L.qpo_gavi_solve(d1, d2, npar, *[_p(a) for a in arrs], _p(w), _p(z0), int(presolve), int(max_pivots),
                          _p(z), _p(zfull), _p(st, ip), _p(pv, ip), _p(basis, bp))
    assert rc == 0
    return dict(z=z, z_full=zfull, status=int(st[0]), pivots=int(pv[0]), basis=basis, z0_projected=z0)


def comp_indices(g, z, w, tol=1e-2):
    L = lib()
    d1, d2 = len(g["l1"]), len(g["l2"])
    arrs = [_f(g[k]) for k in ("M", "N", "o", "l1", "u1", "A", "B", "l2", "u2")]
    mask = np.zeros(d1 + d2, np.int8)
    L.qpo_comp_indices(d1, d2, g["N"].shape[1], *[_p(a) for a in arrs], _p(_f(z)), _p(_f(w)), C.c_double(tol), _p(mask, bp))
    return mask


def halfspace_in(A, l, u, x, tol=1e-6, rl=None, ru=None):
    L = lib()
    A = _f(np.atleast_2d(A))
    m, d = A.shape
    rl = np.zeros(m, np.uint8) if rl is None else np.ascontiguousarray(rl, dtype=np.uint8)
    ru = np.zeros(m, np.uint8) if ru is None else np.ascontiguousarray(ru, dtype=np.uint8)
    return bool(L.qpo_halfspace_in(m, d, _p(A), _p(_f(l)), _p(_f(u)), _p(rl, ubp), _p(ru, ubp), _p(_f(x)), C.c_double(tol)))


def verify_solution(Qd, qd, A, l, u, dec, x, tol=1e-4):
    L = lib()
    Qd = _f(np.atleast_2d(Qd))
    nd, nv = Qd.shape
    A = _f(np.asarray(A, dtype=np.float64).reshape(-1, nv))
    m = A.shape[0]
    dec = np.ascontiguousarray(dec, dtype=np.int32)
    lam = np.zeros(max(m, 1))
    how = np.zeros(1, np.int32)
    act = np.zeros(max(m, 1), np.int8)
    fpiv = np.zeros(1, np.int32)
    sol = L.qpo_verify_solution(nd, nv, m, _p(Qd), _p(_f(qd)), _p(A), _p(_f(l)), _p(_f(u)), _p(dec, ip), _p(_f(x)),
                                C.c_double(tol), _p(lam), _p(how, ip), _p(act, bp), _p(fpiv, ip))
    verify_solution.last_fallback_pivots = int(fpiv[0])
    return bool(sol), lam[:m], int(how[0]), act[:m]


class _Node(C.Structure):
    _fields_ = [("nd", C.c_int32), ("nv", C.c_int32), ("m", C.c_int32),
                ("Qd", dp), ("qd", dp), ("A", dp), ("l", dp), ("u", dp), ("dec", ip)]


class _Gavi(C.Structure):
    _fields_ = [("d1", C.c_int32), ("d2", C.c_int32), ("np", C.c_int32),
                ("M", dp), ("N", dp), ("o", dp), ("l1", dp), ("u1", dp), ("A", dp), ("B", dp), ("l2", dp), ("u2", dp)]


class Level:
    """Column-major copies of a level's data for qpo_level_solve_batched."""

    def __init__(self, nv, views, g, dec, par, max_iters=150, proj=None):
        self.keep = []
        nodes = []
        for (Qd, qd, A, l, u, d) in views:
            Qd = _f(np.atleast_2d(Qd)); A = _f(np.asarray(A, dtype=np.float64).reshape(-1, nv))
            arrs = [Qd, _f(qd), A, _f(l), _f(u), np.ascontiguousarray(d, dtype=np.int32)]
            self.keep.append(arrs)
            nodes.append(_Node(Qd.shape[0], nv, A.shape[0], _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(arrs[5], ip)))
        self.nodes = (_Node * len(nodes))(*nodes)
        ga = [_f(g[k]) for k in ("M", "N", "o", "l1", "u1", "A", "B", "l2", "u2")]
        self.keep.append(ga)
        self.g = _Gavi(len(g["l1"]), len(g["l2"]), g["N"].shape[1], *[_p(a) for a in ga])
        self.dec = np.ascontiguousarray(dec, dtype=np.int32)
        self.par = np.ascontiguousarray(par, dtype=np.int32)
        self.proj = None if proj is None or len(proj) == 0 else np.ascontiguousarray(proj, dtype=np.float64)
        self.nv, self.max_iters = nv, max_iters
        self.lam_total = sum(n.m for n in nodes)

    def solve(self, x_init, threads=1):
        L = lib()
        x_init = np.ascontiguousarray(x_init, dtype=np.float64)
        B = x_init.shape[0]
        x = np.zeros((B, self.nv)); solved = np.zeros(B, np.uint8); iters = np.zeros(B, np.int32); piv = np.zeros(B, np.int32)
        lam = np.zeros((B, max(self.lam_total, 1)))
        nproj = 0 if self.proj is None else self.proj.shape[0]
        L.qpo_level_solve_batched(self.nv, len(self.nodes), self.nodes, C.byref(self.g), _p(self.dec, ip), len(self.dec),
                                  _p(self.par, ip), int(self.max_iters), nproj, _p(self.proj) if nproj else None, B,
                                  _p(x_init), _p(x), _p(solved, ubp), _p(iters, ip), _p(piv, ip), _p(lam), int(threads))
        return dict(x=x, solved=solved.astype(bool), iters=iters, pivots=piv, lam=lam[:, :self.lam_total])
