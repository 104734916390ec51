"""`solve(qpn, init)` and the batched `solve(qpn, inits)` on top of libqpn_cuda.

Mirrors /root/reference/src/requests.jl:1-22 (entry points) and the level loop of
/root/reference/src/algorithm.jl:1-127.  For a level whose players have no children (a flat
Nash game such as four_player_matrix_game with edge_list=[], or the bottom level of any
network) the whole iterate-until-equilibrium loop runs in one kernel launch per batch
(csrc/qpn_level.cuh).  Results keep the reference's NamedTuple fields as dict keys:
solved / x_opt / Sol  or  solved / x_fail / x_opt=None (algorithm.jl:35,116,125).
"""
import numpy as np

from . import assembly
from .engine import Engine, LevelArrays, ResidentLevel
from .examples import SplitMix64


def projection_vectors(net, seed=1):
    """algorithm.jl:10-12: `randn(rng, n)` x num_projections (rng = MersenneTwister(1) in
    requests.jl:21; the documented stand-in generator is used here, see examples.py)."""
    g = SplitMix64(0x9E0 + seed)
    k = net.options.num_projections if net.options.check_for_cycling else 0
    return np.array([[g.randn() for _ in range(net.n_vars)] for _ in range(k)]).reshape(k, net.n_vars)


class BatchedSolver:
    """Holds the GPU-resident levels of one QPNet on one device."""

    def __init__(self, net, engine=None, device=0):
        self.net = net
        self.engine = engine or Engine(device)
        self.proj = projection_vectors(net)
        self._levels = {}

    def resident_level(self, level):
        """The level's players + GAVI uploaded once (no child pieces: bottom level / flat game)."""
        if level not in self._levels:
            net = self.net
            players = net.network_depth_map[level]
            if any(net.network_edges[p] for p in players):
                raise NotImplementedError("levels with children need solution-graph pieces (SURVEY.md 8f-1)")
            g, dec, par = assembly.level_gavi(net, players)
            la = LevelArrays(net.n_vars, [assembly.node_view(net, p) for p in players], g, dec, par,
                             max_iters=net.options.max_iters, proj=self.proj)
            self._levels[level] = ResidentLevel(self.engine, la)
        return self._levels[level]

    def solve_batch(self, inits, out=None, want_lam=False):
        """inits: (B, n_vars).  Returns arrays: solved (B,), x (B, n_vars), iters, pivots."""
        net = self.net
        if net.num_levels() != 1:
            raise NotImplementedError("multi-level networks: batched host recursion is the next row (SURVEY.md 8f-2)")
        return self.resident_level(1).solve(np.ascontiguousarray(inits, dtype=np.float64), out=out, want_lam=want_lam)

    def close(self):
        for lv in self._levels.values():
            lv.release()
        self._levels.clear()


_solvers = {}


def _solver_for(net, device=0):
    key = (id(net), device)
    if key not in _solvers:
        _solvers[key] = BatchedSolver(net, device=device)
    return _solvers[key]


def solve(qpn, x_init=None, device=0):
    """solve(qpn), solve(qpn, x_init) -> one result; solve(qpn, inits::Matrix) -> list of results.

    Julia's `inits::Matrix` is n_vars x B column-major, i.e. a (B, n_vars) C-contiguous array here."""
    x = qpn.default_initialization if x_init is None else np.asarray(x_init, dtype=np.float64)
    single = x.ndim == 1
    ret = _solver_for(qpn, device).solve_batch(np.atleast_2d(x))
    results = []
    for b in range(len(ret["solved"])):
        if ret["solved"][b]:
            results.append(dict(solved=True, x_opt=ret["x"][b].copy(), Sol={}, identified_request=set(), x_alts=[],
                                iters=int(ret["iters"][b]), pivots=int(ret["pivots"][b])))
        else:
            results.append(dict(solved=False, x_fail=ret["x"][b].copy(), x_opt=None,
                                iters=int(ret["iters"][b]), pivots=int(ret["pivots"][b])))
    return results[0] if single else results
