"""ORACLE helper: IEEE fused multiply-add for Python < 3.13 (libm's fma)."""
import ctypes
import ctypes.util
import math

if hasattr(math, "fma"):
    fma = math.fma
else:
    _libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    _libm.fma.restype = ctypes.c_double
    _libm.fma.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]

    def fma(a, b, c):
        return _libm.fma(float(a), float(b), float(c))
