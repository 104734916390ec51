"""ORACLE (test infrastructure): the reference's three example networks as plain
arrays, transcribed by hand from

  /root/reference/examples/simple_bilevel.jl:6-35
  /root/reference/examples/four_player_matrix_game.jl:6-30,118-176
  /root/reference/examples/robust_avoid_simple.jl:1-93

through the rules of add_constraint! / add_qp! / add_edges!
(/root/reference/src/programs.jl:147-201,214-285): A = Jacobian of the
constraint expressions, bounds shifted by the constant term, Q = Hessian of the
cost, q = gradient at 0.

Random problem data: the reference draws it from Julia's MersenneTwister
(dSFMT + ziggurat), which cannot be reproduced without Julia.
It is replaced by the documented generator below (splitmix64 -> U(0,1) ->
Box-Muller), same distributions, fixed seeds.  All indices here are 0-based.
"""
import math
import numpy as np

MASK = (1 << 64) - 1


class SplitMix64:
    """splitmix64 (Steele, Lea, Flood 2014); u01 uses the top 53 bits."""

    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def u01(self):
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def randn(self):
        u1 = 1.0 - self.u01()
        u2 = self.u01()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


class Net:
    """Flat QPNet: what programs.jl:79-116 holds, as dense arrays."""

    def __init__(self, n_vars):
        self.n_vars = n_vars
        self.qps = {}          # id -> dict(Q, q, cons=[cid], vars=[idx])
        self.cons = {}         # cid -> dict(A, l, u)
        self.edges = {}        # id -> sorted list of child ids (minimal adjacency)
        self.reach = {}        # id -> sorted list of reachable ids
        self.depth = {}        # level (1-based) -> sorted list of ids
        self.options = dict(max_iters=150, num_projections=4, exploration_vertices=0,
                            gen_solution_map=False, check_for_cycling=True)
        self.default_init = np.zeros(n_vars)

    def add_constraint(self, A, l, u):
        cid = max(self.cons, default=0) + 1
        self.cons[cid] = dict(A=np.atleast_2d(np.asarray(A, float)),
                              l=np.asarray(l, float), u=np.asarray(u, float))
        return cid

    def add_qp(self, Q, q, cons, vars_):
        pid = max(self.qps, default=0) + 1
        self.qps[pid] = dict(Q=np.asarray(Q, float), q=np.asarray(q, float),
                             cons=list(cons), vars=list(vars_))
        return pid

    def add_edges(self, edge_list):
        """programs.jl:214-285: transitive reduction, reachability, depth map."""
        N = len(self.qps)
        A = np.zeros((N, N), dtype=bool)
        for (i, j) in edge_list:
            assert i != j
            A[i - 1, j - 1] = True
        R = np.zeros((N, N), dtype=bool)
        An = A.copy()
        for _ in range(2, N + 1):
            R |= An
            An = (An.astype(int) @ A.astype(int)) > 0
            for i in range(N):
                assert not An[i, i], "cycle"
                for j in range(N):
                    if A[i, j] and An[i, j]:
                        A[i, j] = False
        deleted, d = set(), 0
        Rd = R.copy()
        while len(deleted) < N:
            nodes = {i for i in range(N) if not Rd[:, i].any()} - deleted
            assert nodes
            d += 1
            self.depth[d] = sorted(i + 1 for i in nodes)
            deleted |= nodes
            remaining = [i for i in range(N) if i not in deleted]
            Rd = R[remaining, :] if remaining else np.zeros((0, N), dtype=bool)
        for i in range(N):
            self.edges[i + 1] = sorted(j + 1 for j in range(N) if A[i, j])
            self.reach[i + 1] = sorted(j + 1 for j in range(N) if R[i, j])

    def decision_inds(self, pid):
        """programs.jl:340-346."""
        inds = set(self.qps[pid]["vars"])
        for j in self.reach[pid]:
            inds |= set(self.qps[j]["vars"])
        return sorted(inds)

    def num_levels(self):
        return len(self.depth)


def simple_bilevel(**opts):
    """vars [w1 w2 x y]; f1 = (y-x)^2 s.t. y>=0 owns y; f2 = |[x;y]-w|^2 owns x; edge 2->1."""
    net = Net(4)
    c = net.add_constraint([[0, 0, 0, 1.0]], [0.0], [math.inf])
    Q1 = np.zeros((4, 4)); Q1[2, 2] = 2; Q1[3, 3] = 2; Q1[2, 3] = Q1[3, 2] = -2
    p1 = net.add_qp(Q1, np.zeros(4), [c], [3])
    Q2 = 2.0 * np.array([[1, 0, -1, 0], [0, 1, 0, -1], [-1, 0, 1, 0], [0, -1, 0, 1.0]])
    p2 = net.add_qp(Q2, np.zeros(4), [], [2])
    net.add_edges([(p2, p1)])
    net.options.update(opts)
    return net


def four_player_constellations(seed=2):
    """Stand-in for `randn(rng,2)` draws of four_player_matrix_game.jl:30."""
    g = SplitMix64(0xF0A4 + seed)
    return {i: {j: np.array([g.randn(), g.randn()]) for j in range(1, 5)} for i in range(1, 5)}


def four_player_matrix_game(edge_list=(), seed=2, **opts):
    """cost_i = |x_i - c_ii|^2 + sum_{j!=i} |x_j - x_i - c_ij|^2, box |x_i|<=5."""
    net = Net(8)
    C = four_player_constellations(seed)
    for i in range(1, 5):
        A = np.zeros((2, 8)); A[0, 2 * (i - 1)] = 1; A[1, 2 * (i - 1) + 1] = 1
        cid = net.add_constraint(A, [-5.0, -5.0], [5.0, 5.0])
        Q = np.zeros((8, 8)); q = np.zeros(8)
        si = slice(2 * (i - 1), 2 * i)
        for j in range(1, 5):
            sj = slice(2 * (j - 1), 2 * j)
            if j == i:
                Q[si, si] += 2 * np.eye(2)
                q[si] += -2 * C[i][j]
            else:
                # d = x_j - x_i - c ; d'd
                Q[sj, sj] += 2 * np.eye(2); Q[si, si] += 2 * np.eye(2)
                Q[si, sj] += -2 * np.eye(2); Q[sj, si] += -2 * np.eye(2)
                q[sj] += -2 * C[i][j]; q[si] += 2 * C[i][j]
        net.add_qp(Q, q, [cid], [2 * (i - 1), 2 * (i - 1) + 1])
    net.add_edges(list(edge_list))
    net.options.update(opts)
    return net


def robust_avoid_data(num_obj=2, num_poly_faces=5, seed=1):
    """Stand-in for the rng draws of robust_avoid_simple.jl:18-28."""
    g = SplitMix64(0x0A01D + seed)
    base = [k * 2 * math.pi / num_poly_faces for k in range(num_poly_faces)]

    def poly():
        noise = [0.15 * g.randn() for _ in range(num_poly_faces)]
        rot = math.pi * g.u01()
        ang = [b + e + rot for b, e in zip(base, noise)]
        A = np.array([[math.cos(a), math.sin(a)] for a in ang])
        return A

    Ae = poly()
    be = (0.2 + 0.8 * g.u01()) * np.ones(num_poly_faces)
    Aos = [poly() for _ in range(num_obj)]
    bos = [(0.2 + 0.8 * g.u01()) * np.ones(num_poly_faces) for _ in range(num_obj)]
    return Ae, be, Aos, bos


def robust_avoid_simple(num_obj=2, num_poly_faces=5, exploration_vertices=10,
                        max_ego_delta=15.0, max_obj_delta=1.0, num_projections=5, seed=1, **opts):
    """Variable order (QPNet(xe,xo,ue,uo,s,eps), column-major flattening):
    xe 0:2, xo 2:2+2k, ue .., uo .., s .., eps .."""
    k, F = num_obj, num_poly_faces
    Ae, be, Aos, bos = robust_avoid_data(k, F, seed)
    ixe = [0, 1]
    ixo = lambda i: [2 + 2 * i, 3 + 2 * i]
    iue = [2 + 2 * k, 3 + 2 * k]
    iuo = lambda i: [4 + 2 * k + 2 * i, 5 + 2 * k + 2 * i]
    is_ = lambda i: [4 + 4 * k + 2 * i, 5 + 4 * k + 2 * i]
    ieps = lambda i: 4 + 6 * k + i
    n = 4 + 7 * k
    net = Net(n)
    s_players, a_players = {}, {}
    for i in range(k):
        A = np.zeros((2 * F, n)); l = np.zeros(2 * F); u = np.full(2 * F, math.inf)
        # Ae*(s - (xe+ue)) + be + eps >= 0
        A[:F, is_(i)] = Ae; A[:F, ixe] = -Ae; A[:F, iue] = -Ae; A[:F, ieps(i)] = 1.0
        l[:F] = -be
        A[F:, is_(i)] = Aos[i]; A[F:, ixo(i)] = -Aos[i]; A[F:, iuo(i)] = -Aos[i]; A[F:, ieps(i)] = 1.0
        l[F:] = -bos[i]
        cid = net.add_constraint(A, l, u)
        q = np.zeros(n); q[ieps(i)] = 1.0
        s_players[i] = net.add_qp(np.zeros((n, n)), q, [cid], is_(i) + [ieps(i)])
    for i in range(k):
        A = np.zeros((2, n)); A[0, iuo(i)[0]] = 1; A[1, iuo(i)[1]] = 1
        cid = net.add_constraint(A, [-max_obj_delta] * 2, [max_obj_delta] * 2)
        q = np.zeros(n); q[ieps(i)] = 1.0
        a_players[i] = net.add_qp(np.zeros((n, n)), q, [cid], iuo(i))
    A = np.zeros((2 + k, n)); A[0, iue[0]] = 1; A[1, iue[1]] = 1
    for i in range(k):
        A[2 + i, ieps(i)] = 1
    cid = net.add_constraint(A, [-max_ego_delta] * 2 + [0.0] * k, [max_ego_delta] * 2 + [math.inf] * k)
    # cost = 0.5 xef'Q xef + xef'q,  Q = diag(0, 0.001), q = [-1, 0], xef = xe + ue
    Q = np.zeros((n, n)); qv = np.zeros(n)
    for a in (ixe[1], iue[1]):
        for b in (ixe[1], iue[1]):
            Q[a, b] = 0.001
    qv[ixe[0]] = -1.0; qv[iue[0]] = -1.0
    ego = net.add_qp(Q, qv, [cid], iue)
    edges = [(ego, a_players[i]) for i in range(k)] + [(a_players[i], s_players[i]) for i in range(k)]
    net.add_edges(edges)
    net.options.update(dict(exploration_vertices=exploration_vertices, num_projections=num_projections))
    net.options.update(opts)
    init = np.zeros(n)
    init[ixe] = [-5.0, 0.0]
    for i in range(k):
        init[ixo(i)] = [3.0 * i, -1.0]
    net.default_init = init
    net.data = dict(Ae=Ae, be=be, Aos=Aos, bos=bos)
    return net
