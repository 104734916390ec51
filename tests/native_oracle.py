"""Test helper: the native network state machine (csrc/net/) built against the C oracle
(oracle/_build/libqpn_net_oracle.so), driven through the product's own marshalling code (qpn_b200.netsolve)."""
import ctypes as C
import os
import subprocess

import numpy as np

import qpn_b200
from qpn_b200 import netsolve

_ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_lib = None


def oracle_net_lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _ORACLE, "-s"])
        _lib = C.CDLL(os.path.join(_ORACLE, "_build", "libqpn_net_oracle.so"))
    return _lib


def oracle_net(qpn, threads=1):
    return netsolve.NetBinding(qpn, oracle_net_lib(), prefix="qpo_net_", threads=threads)


def ra_inits(net, B, seed=0):
    """The perturbed robust_avoid instances of BASELINE.json configs[2]: default init with
    xe, xo moved by N(0, 0.3^2) and ue, uo ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    X = np.tile(net.default_initialization, (B, 1))
    X[:, 0:6] += 0.3 * rng.normal(size=(B, 6))
    X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    return X


def same_result(a, b, sol=False):
    """Two result dicts agree: status, level iteration counts, x bit for bit (and the solution graphs row for row)."""
    if a["solved"] != b["solved"] or list(a["level_iters"]) != list(b["level_iters"]):
        return False
    key = "x_opt" if a["solved"] else "x_fail"
    if not np.array_equal(a[key], b[key]):
        return False
    if sol and a["solved"]:
        for pid, pa in a["Sol"].items():
            pb = b["Sol"].get(pid)
            if pa is None or pb is None:
                if not (pa is None and pb is None):
                    return False
                continue
            if len(pa) != len(pb) or any(p.exact_key != q.exact_key for p, q in zip(pa, pb)):
                return False
    return True
