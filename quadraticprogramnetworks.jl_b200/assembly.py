"""Host assembly of the per-level equilibrium problem: the matrices the kernels receive.

Mirrors /root/reference/src/avi.jl: create_labeled_gavi_from_qp (:205-251), combine_gavis
(:305-377) and the index bookkeeping of solve_qep (:394-404), and the per-node view that
verify_solution reads (/root/reference/src/qp_processing.jl:57-66).  The reference rebuilds
these sparse blocks on every call; here they are built once per (level, child-piece
assignment) and kept resident on the GPU.
"""
import numpy as np

from .model import INF


def node_view(net, pid, pieces=None):
    """(Q[dec,:], q[dec], A, l, u, dec): node `pid` with its own constraint polys followed by
    the chosen solution piece of each child (qp_processing.jl:186-187)."""
    pieces = pieces or {}
    dec = net.decision_inds(pid)
    qp = net.qps[pid]
    polys = [net.constraints[c] for c in qp.constraint_indices] + [pieces[j] for j in net.network_edges[pid]]
    if polys:
        A = np.vstack([p.A for p in polys]); l = np.concatenate([p.l for p in polys]); u = np.concatenate([p.u for p in polys])
    else:
        A, l, u = np.zeros((0, net.n_vars)), np.zeros(0), np.zeros(0)
    return qp.Q[dec, :], qp.q[dec], A, l, u, np.asarray(dec, dtype=np.int32)


def level_gavi(net, players, pieces=None):
    """GAVI of the players' joint KKT system, z = [dec; xi_p...; lambda_p/psi_p...].

    Returns (gavi dict, dec, par).  Block layout per avi.jl:305-377: the first nd rows tie
    each decision variable to the xi of the player(s) owning it, then one stationarity block
    per player (the xi columns carry 0 * -I, avi.jl:244), and A stacks every player's
    constraint rows restricted to the decision columns."""
    pieces = pieces or {}
    players = sorted(players)
    n = net.n_vars
    dec = sorted(set().union(*[set(net.decision_inds(p)) for p in players]))
    par = [i for i in range(n) if i not in set(dec)]
    nd = len(dec)
    dpos = {d: k for k, d in enumerate(dec)}
    views = {p: node_view(net, p, pieces) for p in players}
    xi_dim = {p: len(views[p][5]) for p in players}
    lam_dim = {p: len(views[p][3]) for p in players}
    total_xi, total_lam = sum(xi_dim.values()), sum(lam_dim.values())
    d1, d2 = nd + total_xi, total_lam
    M = np.zeros((d1, d1 + d2)); N = np.zeros((d1, len(par))); o = np.zeros(d1)
    A = np.zeros((d2, d1 + d2)); B = np.zeros((d2, len(par))); l2 = np.zeros(d2); u2 = np.zeros(d2)
    xi_off, lam_off, row = 0, 0, nd
    for p in players:
        Qd, qd, Ap, lp, up, decp = views[p]
        k, m = xi_dim[p], lam_dim[p]
        M[row:row + k, :nd] = Qd[:, dec]
        N[row:row + k, :] = Qd[:, par]
        o[row:row + k] = qd
        M[row:row + k, d1 + lam_off: d1 + lam_off + m] = -Ap[:, decp].T
        for e, dv in enumerate(decp):                       # top rows: sum of the owners' xi = 0
            M[dpos[int(dv)], nd + xi_off + e] = 1.0
        A[lam_off:lam_off + m, :nd] = Ap[:, dec]
        B[lam_off:lam_off + m, :] = Ap[:, par]
        l2[lam_off:lam_off + m] = lp; u2[lam_off:lam_off + m] = up
        row += k; xi_off += k; lam_off += m
    g = dict(M=M, N=N, o=o, l1=np.full(d1, -INF), u1=np.full(d1, INF), A=A, B=B, l2=l2, u2=u2)
    return g, np.asarray(dec, dtype=np.int32), np.asarray(par, dtype=np.int32)
