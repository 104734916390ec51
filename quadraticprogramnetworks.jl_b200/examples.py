"""`setup(:name; kwargs...)` for the reference's three active examples
(/root/reference/src/programs.jl:139-141 dispatching to examples/*.jl).

Random problem data: the reference draws from Julia's MersenneTwister, which cannot be
reproduced outside Julia; the documented stand-in below (splitmix64 -> uniform -> Box-Muller)
draws the same distributions with fixed seeds.  A Julia-built network can be fed to the
engine instead through `QPNet` + `add_constraint` / `add_qp` with its own numbers.
"""
import math

import numpy as np

from .model import INF, Aff, QPNet, Quad, dot, matvec, sumsq

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def rand(self):
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def randn(self):
        u1, u2 = 1.0 - self.rand(), self.rand()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


def setup(name, **kwargs):
    """setup(:name; kwargs...) -> QPNet."""
    name = str(name).lstrip(":")
    try:
        return _SETUPS[name](**kwargs)
    except KeyError:
        raise ValueError(f"unknown example {name!r}; available: {sorted(_SETUPS)}") from None


def setup_simple_bilevel(**kwargs):
    """examples/simple_bilevel.jl:6-35.  variables w1 w2 x y; f1 = (y-x)^2 s.t. y >= 0 owns y;
    f2 = |[x;y] - w|^2 owns x; edge 2 -> 1."""
    net = QPNet(("w", 2), ("x", 1), ("y", 1))
    w, x, y = net.var["w"], net.var["x"][0], net.var["y"][0]
    con_id = net.add_constraint([y], [0.0], [INF])
    qp1 = net.add_qp((y - x) ** 2, [con_id], y)
    qp2 = net.add_qp(sumsq([x - w[0], y - w[1]]), [], x)
    net.add_edges([(qp2, qp1)])
    net.assign_constraint_groups()
    net.set_options(debug_visualize=False, **kwargs)
    net.default_initialization = np.zeros(4)
    return net


def setup_four_player_matrix_game(edge_list=(), seed=2, **kwargs):
    """examples/four_player_matrix_game.jl:6-30,118-176: player i picks x_i in [-5,5]^2 and
    pays |x_i - c_ii|^2 + sum_{j != i} |x_j - x_i - c_ij|^2."""
    net = QPNet(("x1", 2), ("x2", 2), ("x3", 2), ("x4", 2))
    x = {i: net.var[f"x{i}"] for i in range(1, 5)}
    g = SplitMix64(0xF0A4 + seed)
    con = {i: {j: [g.randn(), g.randn()] for j in range(1, 5)} for i in range(1, 5)}
    net.problem_data["constellations"] = con
    for i in range(1, 5):
        con_id = net.add_constraint(x[i], [-5.0, -5.0], [5.0, 5.0])
        cost = 0.0
        for j in range(1, 5):
            if j == i:
                d = [x[i][k] - con[i][j][k] for k in range(2)]
            else:
                d = [x[j][k] - x[i][k] - con[i][j][k] for k in range(2)]
            cost = sumsq(d) + cost
        net.add_qp(cost, [con_id], x[i])
    net.add_edges(list(edge_list))
    net.assign_constraint_groups()
    net.set_options(**kwargs)
    net.default_initialization = np.zeros(8)
    return net


def setup_robust_avoid_simple(num_obj=2, num_poly_faces=5, exploration_vertices=10, max_ego_delta=15.0,
                              max_obj_delta=1.0, num_projections=5, seed=1, max_accel=10.0, **kwargs):
    """examples/robust_avoid_simple.jl:1-93."""
    g = SplitMix64(0x0A01D + seed)
    base = [k * 2 * math.pi / num_poly_faces for k in range(num_poly_faces)]

    def polygon():
        noise = [0.15 * g.randn() for _ in range(num_poly_faces)]
        rot = math.pi * g.rand()
        return np.array([[math.cos(b + e + rot), math.sin(b + e + rot)] for b, e in zip(base, noise)])

    Ae = polygon()
    be = (0.2 + 0.8 * g.rand()) * np.ones(num_poly_faces)
    Aos = [polygon() for _ in range(num_obj)]
    bos = [(0.2 + 0.8 * g.rand()) * np.ones(num_poly_faces) for _ in range(num_obj)]

    net = QPNet(("xe", 2), ("xo", 2 * num_obj), ("ue", 2), ("uo", 2 * num_obj), ("s", 2 * num_obj), ("eps", num_obj))
    xe, ue, eps = net.var["xe"], net.var["ue"], net.var["eps"]
    col = lambda name, i: net.var[name][2 * i:2 * i + 2]          # column i of a 2 x num_obj block
    net.problem_data.update(Ae=Ae, be=be, Ao=Aos, bo=bos)

    s_players, a_players = {}, {}
    for i in range(num_obj):
        s, xo, uo = col("s", i), col("xo", i), col("uo", i)
        rel_e = [s[k] - (xe[k] + ue[k]) for k in range(2)]
        rel_o = [s[k] - (xo[k] + uo[k]) for k in range(2)]
        cons = [r + float(b) + eps[i] for r, b in zip(matvec(Ae, rel_e), be)] + \
               [r + float(b) + eps[i] for r, b in zip(matvec(Aos[i], rel_o), bos[i])]
        con_id = net.add_constraint(cons, [0.0] * len(cons), [INF] * len(cons))
        s_players[i] = net.add_qp(eps[i], [con_id], s, eps[i])
    for i in range(num_obj):
        uo = col("uo", i)
        con_id = net.add_constraint(uo, [-max_obj_delta] * 2, [max_obj_delta] * 2)
        a_players[i] = net.add_qp(eps[i], [con_id], uo)
    con_id = net.add_constraint(list(ue) + list(eps), [-max_ego_delta] * 2 + [0.0] * num_obj,
                                [max_ego_delta] * 2 + [INF] * num_obj)
    Q = np.array([[0.0, 0.0], [0.0, 0.001]])
    q = [-1.0, 0.0]
    xef = [xe[k] + ue[k] for k in range(2)]
    cost = 0.5 * dot(xef, matvec(Q, xef)) + dot(xef, q)           # + 0.5 ue'R ue with R = 0
    ego = net.add_qp(cost, [con_id], ue)

    edges = [(ego, a_players[i]) for i in range(num_obj)] + [(a_players[i], s_players[i]) for i in range(num_obj)]
    net.add_edges(edges)
    net.assign_constraint_groups()
    net.set_options(exploration_vertices=exploration_vertices, num_projections=num_projections, debug_visualize=False, **kwargs)
    init = np.zeros(net.n_vars)
    init[0:2] = [-5.0, 0.0]
    for i in range(num_obj):
        init[2 + 2 * i: 4 + 2 * i] = [3.0 * i, -1.0]
    net.default_initialization = init
    return net


def _randn_matrix(g, r, c, scale=1.0):
    return np.array([g.randn() for _ in range(r * c)]).reshape(r, c) * scale


def setup_monotone_stress(n=256, m=512, seed=5, **kwargs):
    """BASELINE.json configs[4] / SURVEY.md 8d config 5 (no file in the reference: a synthetic stress
    shape): one node,  min 0.5 x'Qx + q'x  s.t.  l <= A x,  Q = G'G + 0.1 I (G iid N(0, 1/n)), rows of A
    iid N(0,1) normalised, l = A xbar - U(0.1, 1) with xbar ~ N(0, I) feasible, q ~ N(0, I).
    Its level AVI is the lifted KKT system of size 2(n + m) (solve_qep form, avi.jl:305-377)."""
    g = SplitMix64(0x5712E55 + seed)
    G = _randn_matrix(g, n, n, 1.0 / math.sqrt(n))
    A = _randn_matrix(g, m, n)
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    xbar = np.array([g.randn() for _ in range(n)])
    l = A @ xbar - np.array([0.1 + 0.9 * g.rand() for _ in range(m)])
    q = np.array([g.randn() for _ in range(n)])
    net = QPNet(("x", n))
    con_id = net.add_constraint([Aff(A[r]) for r in range(m)], l, [INF] * m)
    net.add_qp(Quad(G.T @ G + 0.1 * np.eye(n), q), [con_id], net.var["x"])
    net.add_edges([])
    net.assign_constraint_groups()
    net.set_options(**kwargs)
    net.default_initialization = xbar
    net.problem_data.update(xbar=xbar)
    return net


def setup_synthetic_chain(n=64, levels=3, n_params=8, seed=4, **kwargs):
    """BASELINE.json configs[3] / SURVEY.md 8d config 4 (synthetic): a chain of `levels` nodes with n own
    variables each (+ n_params parameters nobody owns); node k pays 0.5 x_k'(G'G + I) x_k + x_k'C_k [x_parents; p]
    with coupling entries N(0, 0.1^2), subject to n/2 box rows and n/2 random halfspaces through a known
    interior point; edges 1 -> 2 -> ... (node `levels` is the bottom level)."""
    g = SplitMix64(0xC4A12 + seed)
    blocks = [(f"x{k}", n) for k in range(1, levels + 1)] + [("p", n_params)]
    net = QPNet(*blocks)
    nv = net.n_vars
    xint = np.array([g.randn() for _ in range(nv)]) * 0.5
    ids = []
    for k in range(1, levels + 1):
        own = [net.index(v) for v in net.var[f"x{k}"]]
        others = [i for i in range(nv) if i < own[0] or i >= levels * n]          # parents' variables and the parameters
        G = _randn_matrix(g, n, n, 1.0 / math.sqrt(n))
        Q = np.zeros((nv, nv))
        Q[np.ix_(own, own)] = G.T @ G + np.eye(n)
        if others:
            C = _randn_matrix(g, n, len(others), 0.1)
            Q[np.ix_(own, others)] = C
            Q[np.ix_(others, own)] = C.T
        nb = n // 2
        box = [net.var[f"x{k}"][j] for j in range(nb)]
        H = _randn_matrix(g, n - nb, n)
        H /= np.linalg.norm(H, axis=1, keepdims=True)
        half = []
        for r in range(n - nb):
            a = np.zeros(nv); a[own] = H[r]
            half.append(Aff(a))
        lo = [xint[own[j]] - 1.0 for j in range(nb)] + [float(H[r] @ xint[own]) - (0.1 + 0.9 * g.rand()) for r in range(n - nb)]
        up = [xint[own[j]] + 1.0 for j in range(nb)] + [INF] * (n - nb)
        con_id = net.add_constraint(box + half, lo, up)
        ids.append(net.add_qp(Quad(Q, np.zeros(nv)), [con_id], net.var[f"x{k}"]))
    net.add_edges([(ids[k], ids[k + 1]) for k in range(levels - 1)])
    net.assign_constraint_groups()
    net.set_options(**kwargs)
    net.default_initialization = xint
    return net


def robust_avoid_batch(net, batch, seed=0, sigma=0.5):
    """The perturbed instances of BASELINE.json configs[2]: the default initialisation with the
    parameters xe, xo (ego / obstacle positions: nobody owns them) moved by N(0, sigma^2) -- "perturbed obstacles" --
    and the controls ue, uo started at U(-1, 1).  Returns (batch, n_vars)."""
    rng = np.random.default_rng([0xB200, int(seed)])
    nxo = len(net.var["xo"])
    X = np.tile(net.default_initialization, (batch, 1))
    X[:, 0:2 + nxo] += sigma * rng.normal(size=(batch, 2 + nxo))
    lo, hi = 2 + nxo, 2 + nxo + 2 + len(net.var["uo"])
    X[:, lo:hi] = rng.uniform(-1.0, 1.0, (batch, hi - lo))
    return X


_SETUPS = {
    "monotone_stress": setup_monotone_stress,
    "synthetic_chain": setup_synthetic_chain,
    "simple_bilevel": setup_simple_bilevel,
    "four_player_matrix_game": setup_four_player_matrix_game,
    "robust_avoid_simple": setup_robust_avoid_simple,
}
