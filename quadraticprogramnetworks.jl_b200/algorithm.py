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
from .engine import Engine, EngineError, LevelArrays, ResidentLevel
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

    def native_net(self, threads=None):
        """The native state machine for networks with children (qpn_net_*, csrc/net/): built once per net, keeps its
        resident nodes / level GAVIs / pieces across batches."""
        if getattr(self, "_native", None) is None:
            from .netsolve import NetBinding
            self._native = NetBinding(self.net, self.engine.lib, "qpn_net_", handle=self.engine.h)
        if threads:
            self._native.set_option("threads", threads)
        return self._native

    def close(self):
        for lv in self._levels.values():
            lv.release()
        self._levels.clear()
        if getattr(self, "_native", None) is not None:
            self._native.close()
            self._native = None


_solvers = {}


def _fingerprint(net):
    """What a resident solver snapshots: the options and the shape of the net.  The reference lets a user change
    options between solves (set_options!, programs.jl:312-320); a changed fingerprint rebuilds the device copy."""
    import dataclasses
    return (repr(dataclasses.astuple(net.options)), len(net.qps), len(net.constraints), net.n_vars)


def _evict(key):
    s = _solvers.pop(key, None)
    if s is not None:
        try:
            s[1].close()
        except Exception:
            pass


def _solver_for(net, device=0):
    """One BatchedSolver per (net object, device); dropped when the net is garbage collected (weakref.finalize) and
    rebuilt when its options / shape change."""
    import weakref
    key = (id(net), device)
    fp = _fingerprint(net)
    hit = _solvers.get(key)
    if hit is not None and hit[0] != fp:
        _evict(key)
        hit = None
    if hit is None:
        hit = _solvers[key] = (fp, BatchedSolver(net, device=device))
        weakref.finalize(net, _evict, key)
    return hit[1]


def solve(qpn, x_init=None, device=0, workers=1, native=True, threads=None, keep_sol=None):
    """solve(qpn), solve(qpn, x_init) -> one result; solve(qpn, inits::Matrix) -> list of results.

    Julia's `inits::Matrix` is n_vars x B column-major, i.e. a (B, n_vars) C-contiguous array here.
    Networks with children (or with the solution map requested) run on the native batched state machine of
    libqpn_cuda (`qpn_net_solve_batched`; `threads` host threads drive the batch); `keep_sol` (default: single
    instances only) also returns the solution graphs `Sol`.  native=False runs the Python host mirror of the same
    logic instead (one recursion per instance; `workers` > 1 shards it over that many processes, workers.py)."""
    x = qpn.default_initialization if x_init is None else np.asarray(x_init, dtype=np.float64)
    single = x.ndim == 1
    solver = _solver_for(qpn, device)
    if (qpn.num_levels() != 1 or qpn.options.gen_solution_map) and native:
        outs = solver.native_net(threads).solve(np.atleast_2d(x), keep_sol=single if keep_sol is None else keep_sol)
        return outs[0] if single else outs
    if qpn.num_levels() != 1 or qpn.options.gen_solution_map:
        # networks with children (or with the solution map requested): the per-instance recursion of solve_base!,
        # every numeric step on the device.  A batch runs one recursion per instance against a BatchingEngine,
        # which regroups the instances' pending device calls into one launch per (kind, shared data) (batching.py).
        X = np.atleast_2d(x)
        if len(X) == 1:
            outs = [NetSolver(qpn, solver.engine).solve(X[0])]
        elif workers > 1:
            from .workers import solve_multilevel_workers
            outs = solve_multilevel_workers(qpn, X, workers, engine=solver.engine)
        else:
            outs = solve_multilevel_batch(qpn, X, solver.engine)
        return outs[0] if single else outs
    ret = solver.solve_batch(np.atleast_2d(x))
    results = []
    for b in range(len(ret["solved"])):
        if ret["solved"][b]:
            results.append(dict(solved=True, x_opt=ret["x"][b].copy(), Sol={}, identified_request=set(), x_alts=[],
                                iters=int(ret["iters"][b]), pivots=int(ret["pivots"][b])))
        else:
            results.append(dict(solved=False, x_fail=ret["x"][b].copy(), x_opt=None,
                                iters=int(ret["iters"][b]), pivots=int(ret["pivots"][b])))
    return results[0] if single else results


def solve_multilevel_batch(qpn, X, engine, chunk=256, stats=None, pieces=None, memo=None):
    """solve(qpn, inits) for a network with children: one NetSolver per instance (its own iterate cache and cycle
    detection), all of them sharing the memoised pieces and driving the device through one BatchingEngine.
    pieces / memo: the piece and predicate memos, when the caller keeps them across batches (workers.py)."""
    from .batching import BatchingEngine
    pieces = {} if pieces is None else pieces
    memo = {} if memo is None else memo
    outs = []
    for lo in range(0, len(X), chunk):
        be = BatchingEngine(engine)
        jobs = [(lambda xi=xi: NetSolver(qpn, be, piece_cache=pieces, lp_memo=memo).solve(xi)) for xi in X[lo:lo + chunk]]
        outs += be.run(jobs)
        if stats is not None:
            stats["rounds"] = stats.get("rounds", 0) + be.rounds
            stats["device_calls"] = stats.get("device_calls", 0) + be.device_calls
            stats["requests"] = stats.get("requests", 0) + be.requests
    return outs


def flatten(qpn):
    """programs.jl:117-124: the same players with every edge removed (a flat Nash game)."""
    import copy
    flat = copy.deepcopy(qpn)
    flat.network_edges, flat.reachable_nodes, flat.network_depth_map = {}, {}, {}
    flat._dec_cache = {}
    flat.add_edges([])
    return flat


def get_flat_initialization(qpn, x0=None, device=0):
    """programs.jl:126-131: the equilibrium of the flattened net, as a start for the real one."""
    flat = flatten(qpn)
    flat.options.gen_solution_map = False
    ret = solve(flat, np.zeros(qpn.n_vars) if x0 is None else x0, device=device)
    return ret["x_opt"]


# ==============================================================================================
# Multi-level networks: host recursion of solve_base! with the numeric steps on the device
# ==============================================================================================
import itertools  # noqa: E402
import math  # noqa: E402

from . import polyhedra as ph  # noqa: E402
from . import solgraph  # noqa: E402


class SolveError(RuntimeError):
    """The reference raises inside solve_base! and turns it into solved=false (algorithm.jl:120-126)."""


def _julia_product(ranges):
    """Iterators.product order: the FIRST iterator varies fastest (qp_processing.jl:169)."""
    for tup in itertools.product(*reversed(ranges)):
        yield tuple(reversed(tup))


def intersection_leaves(unions, red_lengths, x, lp, engine=None):
    """IntersectionRoot (intersection.jl:55-151): all non-empty intersections of one piece per union
    that contain x in their closure, skipping the combinations made of complements only.

    The membership test `central_point in closure(poly)` (intersection.jl:74,82) is a conjunction over
    the pieces of a branch, so it is evaluated once per piece for the whole tree by the batched
    half-space kernel (qpn_halfspace_in_batched); only survivors reach the emptiness LP."""
    n = len(unions)
    full = [len(u) for u in unions]
    out = []
    inside = {}
    if engine is not None and hasattr(engine, "halfspace_in"):
        flat = [p for u in unions for p in u if len(p)]
        if flat:
            got = engine.halfspace_in([(p.A, p.l, p.u) for p in flat], np.asarray(x, dtype=np.float64)[None, :], tol=1e-6)[0]
            inside = {id(p): bool(v) for p, v in zip(flat, got)}

    def alive(piece, poly):
        member = inside[id(piece)] if id(piece) in inside else ph.contains(piece, x, closed=True)
        return member and not ph.isempty(poly, lp)

    def rec(depth, poly, idx):
        if depth == n:
            if all(i >= f - r for i, f, r in zip(idx, full, red_lengths)):
                return                                    # the all-complements "red zone" (intersection.jl:123)
            out.append(poly)
            return
        for k, piece in enumerate(unions[depth]):
            cur = piece if poly is None else ph.intersect(piece, poly)
            if not alive(piece, cur):                    # the parent already contains x: only the new piece is tested
                continue
            rec(depth + 1, cur, idx + [k])

    rec(0, None, [])
    return out


class NetSolver:
    """solve(qpn, x_init) for any QPNet: algorithm.jl:1-127 + qp_processing.jl:151-291 on the host,
    with verify_solution / solve_qep / comp_indices / every LP on the device engine."""

    def __init__(self, net, engine, piece_cache=None, lp_memo=None):
        self.net, self.engine = net, engine
        self.lp = ph.LPSolver(engine)
        if lp_memo is not None:
            self.lp.memo = lp_memo     # shared by the instance threads of a batch (batching.py)
        self.proj = projection_vectors(net)
        self.iterate_cache = {}
        # (node, constraint polys, K) -> local piece / its projection; lives across instances
        self.piece_cache = {} if piece_cache is None else piece_cache

    # ---- qp_processing.jl:57-149 on the device -----------------------------------------------
    def verify(self, pid, polys, dec, x):
        net = self.net
        qp = net.qps[pid]
        if polys:
            A = np.vstack([p.A for p in polys]); l = np.concatenate([p.l for p in polys]); u = np.concatenate([p.u for p in polys])
        else:
            A, l, u = np.zeros((0, net.n_vars)), np.zeros(0), np.zeros(0)
        if net.options.check_convexity:                 # qp_processing.jl:69 (raises when the node is not convex)
            from .qp import check_qp_convexity
            try:
                check_qp_convexity(self.engine, qp.Q, A, l, u, dec, pid)
            except ValueError as err:
                raise SolveError(str(err)) from None
        sol, lam, how, act = self.engine.verify_solution((qp.Q[dec, :], qp.q[dec], A, l, u, np.asarray(dec, np.int32)), x[None, :])
        return bool(sol[0]), lam[0]

    # ---- avi.jl:382-444 on the device ------------------------------------------------------------
    def solve_qep(self, players, x, pieces):
        g, dec, par = assembly.level_gavi(self.net, players, pieces)
        z0 = np.zeros((1, g["M"].shape[1])); z0[0, :len(dec)] = x[dec]
        ret = self.engine.gavi_solve(g, x[par][None, :], z0)
        if int(ret["status"][0]) != 1:
            raise SolveError("AVI solve error. This might be because one of the qps is unbounded or ill-conditioned.")
        x_opt = x.copy()
        x_opt[dec] = ret["z"][0, :len(dec)]
        return x_opt

    def solution_pieces(self, pid, polys, dec, x, lam):
        gen = solgraph.process_solution_graph(self.net, pid, polys, dec, x, lam, self.engine, self.lp,
                                              exploration_vertices=self.net.options.exploration_vertices, cache=self.piece_cache)
        return gen.collect()

    # ---- qp_processing.jl:151-241 ------------------------------------------------------------------
    # process_qp in two phases: a level first verifies every player against every combination of child pieces
    # (`verify_phase`), then builds the graph of every player that verified (`graph_phase`) -- as the reference does per
    # player, including for a level that is NOT at an equilibrium, where the graphs are thrown away (algorithm.jl:47-52,
    # 68-101) but an error raised while building one, or a failed combine, still ends the solve (algorithm.jl:56-63,120-126).
    def verify_phase(self, pid, x, S):
        net = self.net
        base = [net.constraints[c] for c in net.qps[pid].constraint_indices]
        dec = net.decision_inds(pid)
        children = sorted(net.network_edges[pid])
        if children:
            if any(len(S[j]) < 1 for j in children):
                raise SolveError("Solution graphs were not properly populated.")
            combos = []
            for combo in _julia_product([range(len(S[j])) for j in children]):
                pieces = [S[j][ji] for j, ji in zip(children, combo)]
                polys = base + pieces
                ok, lam = self.verify(pid, polys, dec, x)
                combos.append((ok, combo, pieces, polys, lam))
            for ok, combo, _, _, _ in combos:
                if not ok:
                    return dict(solution=False, subpiece_assignments=dict(zip(children, combo)))
            return dict(solution=True, combos=combos, dec=dec)
        ok, lam = self.verify(pid, base, dec, x)
        if not ok:
            return dict(solution=False, subpiece_assignments={})
        return dict(solution=True, combos=None, base=base, lam=lam, dec=dec)

    def graph_phase(self, pid, x, vr):
        """The solution graph of a player that verified: dict(S=..., failed=...)."""
        net = self.net
        gen = (pid not in net.network_depth_map[1]) or net.options.gen_solution_map
        if not gen:
            return dict(S=None, failed=False)
        dec = vr["dec"]
        if vr["combos"] is not None:
            graphs = [(pieces, ph.remove_subsets(self.solution_pieces(pid, polys, dec, x, lam), self.lp))
                      for _, _, pieces, polys, lam in vr["combos"]]
            try:
                return dict(S=self.combine(graphs, x), failed=False)
            except SolveError:
                return dict(S=None, failed=True)
        S_out = self.solution_pieces(pid, vr["base"], dec, x, vr["lam"])
        if len(S_out) == 0:
            raise SolveError("This shouldn't happen. Solution graph is empty.")
        return dict(S=S_out, failed=False)

    def process_qp(self, pid, x, S):
        """qp_processing.jl:151-241 for one player (both phases)."""
        vr = self.verify_phase(pid, x, S)
        if not vr["solution"]:
            return dict(solution=False, failed=False, subpiece_assignments=vr["subpiece_assignments"])
        return dict(solution=True, **self.graph_phase(pid, x, vr))

    # ---- qp_processing.jl:243-291 ---------------------------------------------------------------------
    def combine(self, solgraphs, x):
        regions = [ph.intersect(*r) for r, _ in solgraphs]
        solutions = [s for _, s in solgraphs]
        if len(solutions) == 0:
            raise SolveError("No solutions to combine")
        if len(solutions) == 1:
            return list(solutions[0])
        complements = [ph.complement(r) for r in regions]
        combined = [list(s) + rc for s, rc in zip(solutions, complements)]
        widths = [len(c) for c in combined]
        if len(widths) > 3 and sum(widths) > 20:
            raise SolveError("Too many solutions to combine.")
        return intersection_leaves(combined, [len(c) for c in complements], x, self.lp, self.engine)

    # ---- algorithm.jl:1-127 -------------------------------------------------------------------------------
    def solve(self, x_init):
        self.iterate_cache = {}
        self.level_iters = {}                               # level -> loop passes of solve_base, summed over its calls
        ret = self.solve_base(np.asarray(x_init, dtype=np.float64), 1)
        ret["level_iters"] = [self.level_iters.get(lv, 0) for lv in range(1, self.net.num_levels() + 1)]
        return ret

    def solve_base(self, x_init, level):
        net, opt = self.net, self.net.options
        x = x_init.copy()
        try:
            for _ in range(opt.max_iters):
                if hasattr(self, "level_iters"):
                    self.level_iters[level] = self.level_iters.get(level, 0) + 1
                if opt.check_for_cycling:
                    if opt.num_projections == 0:
                        raise SolveError("Cycling check requested, but num_projections == 0.")
                    pv = self.proj @ x
                    cache = self.iterate_cache.setdefault(level, [])
                    for prev in cache:
                        if np.linalg.norm(pv - prev) <= 1.4901161193847656e-8 * max(np.linalg.norm(pv), np.linalg.norm(prev)):
                            raise SolveError("Cycling detected (noticed solution iterate returned to a previous value).")
                    cache.append(pv)
                if level < net.num_levels():
                    low = self.solve_base(x, level + 1)
                    if not low["solved"]:
                        return dict(solved=False, x_fail=x, x_opt=None)
                    S, x = low["Sol"], low["x_opt"]
                else:
                    S = {}
                players = sorted(net.network_depth_map[level])
                children = sorted(set().union(*[set(net.network_edges[i]) for i in players]))
                vrs = [self.verify_phase(pid, x, S) for pid in players]
                equilibrium = all(vr["solution"] for vr in vrs)
                assignments = {i: S[i][0] for i in children}
                if equilibrium:
                    results = [self.graph_phase(pid, x, vr) for pid, vr in zip(players, vrs)]
                    if any(r["failed"] for r in results):
                        return dict(solved=False, x_fail=x, x_opt=None)
                    for pid, r in zip(players, results):
                        S[pid] = ph.remove_subsets(r["S"], self.lp) if (r["S"] is not None and self._removes(level)) else r["S"]
                else:
                    # the reference has built (and will discard) the graph of every player that did verify: what raises or
                    # fails there ends the solve (solve_base's catch-all turns a raise into solved=false)
                    for pid, vr in zip(players, vrs):
                        if vr["solution"] and self.graph_phase(pid, x, vr)["failed"]:
                            return dict(solved=False, x_fail=x, x_opt=None)
                if (not equilibrium) and level < net.num_levels():
                    for vr in vrs:
                        if not vr["solution"]:
                            for child, sub in vr["subpiece_assignments"].items():
                                assignments[child] = S[child][sub]
                if not equilibrium:
                    xnew = self.solve_qep(players, x, assignments)
                    if np.linalg.norm(xnew - x) < 1e-4:
                        raise SolveError("Detected disagreement in solution status between qp solution processer and equilibrium solver.")
                    x = xnew
                    continue
                if level == 1:
                    self.iterate_cache = {}
                return dict(solved=True, x_opt=x, Sol=S, identified_request=set(), x_alts=[])
            raise SolveError("Can't find solution")
        except EngineError:
            raise                                           # a CUDA / library error is not an outcome of the instance
        except Exception as err:                            # noqa: BLE001 -- algorithm.jl:120-126: `catch err` -> solved=false
            self.iterate_cache = {}
            return dict(solved=False, x_fail=x, x_opt=None, error=str(err) or type(err).__name__)

    def _removes(self, level):
        lv = self.net.options.levels_to_remove_subsets
        return True if lv is None else (level in lv)
