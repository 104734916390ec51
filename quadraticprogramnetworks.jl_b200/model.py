"""Host-side model of a Quadratic Program Network: the data `setup(:name)` produces in the
reference, as plain arrays.

Mirrors /root/reference/src/programs.jl (QPNet :79-116, add_constraint! :147-170, add_qp!
:172-201, add_edges! :214-285, decision_inds :340-346, QPNetOptions :61-77) and the Slice /
Poly normalisation of /root/reference/src/sets.jl:68-92,151-158.  The reference derives
A, Q, q from Symbolics expressions; here a small affine / quadratic expression algebra
(`Aff`, `Quad`) does the same job without a CAS, so example files read like the originals.
"""
import math
from dataclasses import dataclass, field

import numpy as np

INF = math.inf


# ----------------------------------------------------------------------------------------
# expressions
# ----------------------------------------------------------------------------------------
class Aff:
    """a'x + c over the network's variable vector."""

    __array_priority__ = 1000

    def __init__(self, a, c=0.0):
        self.a = np.asarray(a, dtype=float)
        self.c = float(c)

    def _lift(self, o):
        return o if isinstance(o, Aff) else Aff(np.zeros_like(self.a), float(o))

    def __add__(self, o):
        if isinstance(o, Quad):
            return o + self
        o = self._lift(o)
        return Aff(self.a + o.a, self.c + o.c)

    __radd__ = __add__

    def __neg__(self):
        return Aff(-self.a, -self.c)

    def __sub__(self, o):
        return self + (-o if isinstance(o, (Aff, Quad)) else -float(o))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Aff):
            Q = np.outer(self.a, o.a) + np.outer(o.a, self.a)
            return Quad(Q, self.c * o.a + o.c * self.a, self.c * o.c)
        return Aff(self.a * float(o), self.c * float(o))

    __rmul__ = __mul__

    def __pow__(self, p):
        assert p == 2
        return self * self


class Quad:
    """0.5 x'Qx + q'x + k  (Q symmetric): what add_qp! extracts, programs.jl:173-187."""

    def __init__(self, Q, q, k=0.0):
        self.Q, self.q, self.k = np.asarray(Q, float), np.asarray(q, float), float(k)

    def __add__(self, o):
        if isinstance(o, Quad):
            return Quad(self.Q + o.Q, self.q + o.q, self.k + o.k)
        if isinstance(o, Aff):
            return Quad(self.Q, self.q + o.a, self.k + o.c)
        return Quad(self.Q, self.q, self.k + float(o))

    __radd__ = __add__

    def __neg__(self):
        return Quad(-self.Q, -self.q, -self.k)

    def __sub__(self, o):
        return self + (-o if isinstance(o, (Aff, Quad)) else -float(o))

    def __mul__(self, s):
        return Quad(self.Q * float(s), self.q * float(s), self.k * float(s))

    __rmul__ = __mul__


def sumsq(exprs):
    """sum_i e_i^2 for affine e_i  (the d'd terms of the examples)."""
    out = 0.0
    for e in exprs:
        out = e * e + out
    return out


def dot(u, v):
    out = 0.0
    for a, b in zip(u, v):
        out = (a * b) + out
    return out


def matvec(M, v):
    M = np.asarray(M, float)
    return [sum((float(M[i, j]) * v[j] for j in range(len(v))), start=Aff(np.zeros_like(v[0].a))) for i in range(M.shape[0])]


# ----------------------------------------------------------------------------------------
# polyhedra as stacked normalised slices
# ----------------------------------------------------------------------------------------
def normalize_slice(a, l, u, rl=False, ru=False, tol=1e-8):
    """sets.jl:76-89: drop tiny entries, scale so the leading non-zero is +1 (flipping the
    bounds when it was negative).  rl / ru: True = strict '<'."""
    a = np.array(a, dtype=float)
    a[np.abs(a) <= tol] = 0.0
    if math.sqrt(float(a @ a)) <= tol:
        return np.zeros_like(a), float(l), float(u), rl, ru
    lead = a[np.flatnonzero(a)[0]]
    n = abs(lead)
    if lead >= 0:
        return a / n, l / n, u / n, rl, ru
    return -a / n, -u / n, -l / n, ru, rl


def _r5(v):
    return (v if math.isinf(v) else round(v, 5)) + 0.0


class Poly:
    """BasicPoly (sets.jl:123-125): rows A, bounds l <= Ax <= u, open/closed flags.  Equal
    slices (5-digit rounding, sets.jl:104-112) are stored once, in first-seen order."""

    def __init__(self, A, l, u, rl=None, ru=None, normalize=True):
        A = np.atleast_2d(np.asarray(A, dtype=float))
        l, u = np.asarray(l, dtype=float).reshape(-1), np.asarray(u, dtype=float).reshape(-1)
        m = len(l)
        d = A.shape[1]
        A = A.reshape(m, d)
        rl = np.zeros(m, bool) if rl is None else np.asarray(rl, bool).reshape(-1)
        ru = np.zeros(m, bool) if ru is None else np.asarray(ru, bool).reshape(-1)
        if normalize and m:
            # every row through normalize_slice (sets.jl:76-89), all rows at once
            A = A.copy()
            A[np.abs(A) <= 1e-8] = 0.0
            zero = np.sqrt(np.einsum("ij,ij->i", A, A)) <= 1e-8
            A[zero] = 0.0
            first = (A != 0.0).argmax(axis=1)
            lead = A[np.arange(m), first]
            n = np.abs(lead)
            n[zero] = 1.0
            pos = (lead >= 0) | zero
            A = A / n[:, None]
            A[~pos] = -A[~pos]
            l, u = np.where(pos, l / n, -u / n), np.where(pos, u / n, -l / n)
            rl, ru = np.where(pos, rl, ru), np.where(pos, ru, rl)
        elif m:
            A = A.copy()
        # equal slices (5-digit rounding, sets.jl:104-112) are stored once, in first-seen order
        K = np.empty((m, d + 4))
        np.round(A, 5, out=K[:, :d])
        with np.errstate(invalid="ignore"):
            K[:, d] = np.where(np.isinf(l), l, np.round(l, 5))
            K[:, d + 1] = np.where(np.isinf(u), u, np.round(u, 5))
        K[:, d + 2] = rl
        K[:, d + 3] = ru
        K += 0.0                                           # -0.0 -> +0.0: equal keys have equal bytes
        keys = [row.tobytes() for row in K]
        if m > 1 and len(set(keys)) < m:
            seen, keep = set(), []
            for i, key in enumerate(keys):
                if key not in seen:
                    seen.add(key); keep.append(i)
            A, l, u, rl, ru = A[keep], l[keep], u[keep], rl[keep], ru[keep]
            keys = [keys[i] for i in keep]
        self.A = np.ascontiguousarray(A, dtype=float).reshape(len(l), d)
        self.l = np.array(l, dtype=float)
        self.u = np.array(u, dtype=float)
        self.rl = np.array(rl, dtype=bool)
        self.ru = np.array(ru, dtype=bool)
        self._K = keys
        self._keyset = None
        self._exact = None

    @property
    def exact_key(self):
        """The polyhedron as stored, row order and every bit included: the key of every memo whose value depends on
        the rows themselves (node GAVIs, pieces, LP results) -- the 5-digit set equality below is the reference's
        notion of equal SETS, not of equal data."""
        if self._exact is None:
            self._exact = (self.A.shape, self.A.tobytes(), self.l.tobytes(), self.u.tobytes(), self.rl.tobytes(), self.ru.tobytes())
        return self._exact

    @property
    def _keys(self):
        """The set of slice keys (built on first use: most polyhedra are temporaries that are never compared)."""
        if self._keyset is None:
            self._keyset = frozenset(self._K)
        return self._keyset

    def __len__(self):
        return len(self.l)

    @property
    def dim(self):
        return self.A.shape[1]

    def __eq__(self, other):                      # sets.jl:141-146
        return isinstance(other, Poly) and self._keys == other._keys

    def __hash__(self):
        return hash(self._keys)

    def closure(self):                            # sets.jl:364-366
        return Poly(self.A, self.l, self.u, normalize=False)


@dataclass
class QP:
    Q: np.ndarray
    q: np.ndarray
    k: float
    constraint_indices: list
    var_indices: list                              # 0-based


@dataclass
class QPNetOptions:
    """programs.jl:61-77 (dead options kept so set_options! accepts the same names)."""
    shared_variable_mode: str = "SHARED_DUAL"
    max_iters: int = 150
    tol: float = 1e-4
    high_dimension: bool = False
    high_dimension_max_iters: int = 10
    num_projections: int = 4
    make_requests: bool = False
    exploration_vertices: int = 0
    try_hull: bool = False
    debug_visualize: bool = False
    gen_solution_map: bool = False
    levels_to_remove_subsets: object = None        # None = every level (NaturalNumbers)
    check_convexity: bool = False
    check_for_cycling: bool = True
    perturb_to_continue: bool = True


class QPNet:
    """QPNet(vars...) of programs.jl:94-116.  Variables are declared as (name, count) blocks
    and addressed as `net.var[name][k]` (an `Aff`)."""

    def __init__(self, *blocks):
        self.names, self.var = [], {}
        n = sum(c for _, c in blocks)
        off = 0
        for name, c in blocks:
            self.var[name] = [Aff(np.eye(n)[off + k]) for k in range(c)]
            self.names += [f"{name}[{k + 1}]" for k in range(c)]
            off += c
        self.n_vars = n
        self.qps, self.constraints = {}, {}
        self.network_edges, self.reachable_nodes, self.network_depth_map = {}, {}, {}
        self.options = QPNetOptions()
        self.default_initialization = np.zeros(n)
        self.problem_data = {}
        self._dec_cache = {}

    def index(self, v):
        """0-based index of a variable handle."""
        nz = np.flatnonzero(v.a)
        assert len(nz) == 1 and v.c == 0.0
        return int(nz[0])

    # programs.jl:147-170
    def add_constraint(self, cons, lb, ub, tol=1e-8):
        cons = list(cons)
        assert len(cons) == len(lb) == len(ub)
        A = np.array([c.a for c in cons]).reshape(len(cons), self.n_vars)
        A[np.abs(A) <= tol] = 0.0
        vals = np.array([c.c for c in cons])
        cid = max(self.constraints, default=0) + 1
        self.constraints[cid] = Poly(A, np.asarray(lb, float) - vals, np.asarray(ub, float) - vals)
        return cid

    # programs.jl:172-201
    def add_qp(self, cost, con_inds, *private_vars, tol=1e-8):
        if isinstance(cost, Aff):
            cost = Quad(np.zeros((self.n_vars, self.n_vars)), cost.a, cost.c)
        Q = cost.Q.copy()
        Q[np.abs(Q) <= tol] = 0.0
        var_inds = []
        for v in private_vars:
            var_inds += [self.index(x) for x in (v if isinstance(v, (list, tuple)) else [v])]
        pid = max(self.qps, default=0) + 1
        self.qps[pid] = QP(Q, cost.q.copy(), cost.k, list(con_inds), var_inds)
        self._dec_cache.clear()
        return pid

    # programs.jl:214-285
    def add_edges(self, edge_list):
        N = len(self.qps)
        A = np.zeros((N, N), dtype=bool)
        for (i, j) in edge_list:
            if i == j:
                raise ValueError(f"Cannot have self edges. (In this case, node {i} -> {i}).")
            A[i - 1, j - 1] = True
        R = np.zeros((N, N), dtype=bool)
        An = A.copy()
        for n in range(2, N + 1):
            R |= An
            An = (An.astype(np.int64) @ A.astype(np.int64)) > 0
            if np.diag(An).any():
                raise ValueError("Cycle detected.")
            A &= ~An                                  # an edge with an alternate longer path is redundant
        depth, deleted = 0, set()
        Rd = R
        while len(deleted) < N:
            nodes = {i for i in range(N) if not Rd[:, i].any()} - deleted
            if not nodes:
                raise ValueError("Something appears wrong with the graph structure.")
            depth += 1
            self.network_depth_map[depth] = sorted(i + 1 for i in nodes)
            deleted |= nodes
            remaining = [i for i in range(N) if i not in deleted]
            Rd = R[remaining, :] if remaining else np.zeros((0, N), dtype=bool)
        for i in range(N):
            self.network_edges[i + 1] = sorted(int(j) + 1 for j in np.flatnonzero(A[i]))
            self.reachable_nodes[i + 1] = sorted(int(j) + 1 for j in np.flatnonzero(R[i]))
        self._dec_cache.clear()

    def assign_constraint_groups(self, group_map=None):
        """programs.jl:293-310 -- multiplier groups only matter in the dead MIN_NORM mode."""
        self.group_map = group_map or {}

    # programs.jl:312-320
    def set_options(self, **kwargs):
        for k, v in kwargs.items():
            if hasattr(self.options, k):
                setattr(self.options, k, v)
            else:
                import warnings
                warnings.warn(f"Invalid option name {k} with value {v}, skipping")

    def num_levels(self):
        return len(self.network_depth_map)

    # programs.jl:340-346
    def decision_inds(self, pid):
        if pid not in self._dec_cache:
            inds = set(self.qps[pid].var_indices)
            for j in self.reachable_nodes[pid]:
                inds |= set(self.qps[j].var_indices)
            self._dec_cache[pid] = sorted(inds)
        return self._dec_cache[pid]
