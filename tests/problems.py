"""Seeded problem generators shared by the CPU and GPU test suites."""
import math

import numpy as np

from oracle import examples, qpn_ref

INF = math.inf


def qp_gavi(Q, c, A, l, u):
    """KKT system of  min 0.5 x'Qx + c'x  s.t. l <= Ax <= u  as a GAVI (avi.jl:447-475 shape)."""
    n, m = len(c), len(l)
    return dict(M=np.hstack([Q, -A.T]), N=np.zeros((n, 0)), o=np.asarray(c, float), l1=np.full(n, -INF), u1=np.full(n, INF),
                A=np.hstack([A, np.zeros((m, m))]), B=np.zeros((m, 0)), l2=np.asarray(l, float), u2=np.asarray(u, float))


def random_qp(rng, kind, n=None, m=None):
    """kind 0: LP in a box, 1: strictly convex QP with two-sided rows, 2: rank-1 QP in a box."""
    n = n or int(rng.integers(2, 6))
    m = m or int(rng.integers(n + 1, 3 * n + 3))
    A = rng.normal(size=(m, n))
    xbar = rng.normal(size=n)
    l = A @ xbar - rng.uniform(0.1, 1, m)
    u = np.full(m, INF)
    Q = np.zeros((n, n))
    if kind == 1:
        u = A @ xbar + rng.uniform(0.1, 1, m)
        G = rng.normal(size=(n, n)); Q = G.T @ G
    elif kind == 2:
        G = rng.normal(size=(1, n)); Q = G.T @ G
    c = rng.normal(size=n)
    if kind != 1:
        A = np.vstack([A, np.eye(n)]); l = np.concatenate([l, xbar - 3]); u = np.concatenate([u, xbar + 3])
    z0 = np.concatenate([2 * rng.normal(size=n), np.zeros(len(l))])
    return Q, c, A, l, u, z0


def fp_avi():
    """The level-1 AVI of the four-player Nash game (no parameters: q = o for every instance)."""
    net = examples.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    return net, g, qpn_ref.convert(g), dec, par


def fp_starts(rng, B):
    """z0s of solve_gavi for inits ~ U(-5,5)^8 (four_player_matrix_game.jl:123-125 box)."""
    X = rng.uniform(-5, 5, (B, 8))
    z0 = np.zeros((B, 32)); z0[:, :8] = X; z0[:, 24:] = X
    return X, z0


def ra_inits(rng, B, net=None):
    """robust_avoid_simple perturbed inits."""
    net = net or examples.robust_avoid_simple()
    X = np.tile(net.default_init, (B, 1))
    X[:, 0:6] += 0.5 * rng.normal(size=(B, 6))
    X[:, 6:12] = rng.uniform(-1, 1, (B, 6))
    return net, X


def dense_to_csc(M, base=1):
    """Julia-style CSC (1-based by default) of a dense matrix."""
    M = np.asarray(M)
    colptr, rowval, nzval = [base], [], []
    for j in range(M.shape[1]):
        for i in range(M.shape[0]):
            if M[i, j] != 0.0:
                rowval.append(i + base); nzval.append(M[i, j])
        colptr.append(len(rowval) + base)
    return np.array(colptr, np.int32), np.array(rowval, np.int32), np.array(nzval, float)
