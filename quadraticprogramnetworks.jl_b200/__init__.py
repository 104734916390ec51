"""qpn_b200: B200-native equilibrium engine for Quadratic Program Networks.

Host-side mirror of the reference's API for the numeric hot path
(`setup(:name)` -> `solve(qpn, init)`, plus the batched `solve(qpn, inits)`), driving
libqpn_cuda through its C ABI (include/qpn_cuda.h).  Import as `qpn_b200` (the shim
at the repo root) -- the directory name carries a dot and is not importable directly.
"""
from .engine import Engine, EngineError, GaviArrays, NodeArrays, LevelArrays, ResidentLevel, load_library, LIB_PATH  # noqa: F401
from .model import QPNet, QP, Poly, Aff, Quad, QPNetOptions, sumsq, dot, matvec, INF  # noqa: F401,E402
from .examples import setup  # noqa: F401,E402
from .algorithm import solve, BatchedSolver, NetSolver, projection_vectors, flatten, get_flat_initialization, solve_multilevel_batch  # noqa: F401,E402
from .workers import MultilevelPool, solve_multilevel_workers  # noqa: F401,E402
from .export import export_net, load_net  # noqa: F401,E402
from .qp import solve_qp, check_qp_convexity, implicit_bounds  # noqa: F401,E402
from . import assembly, examples, model, algorithm, engine, sharding, polyhedra, solgraph, export, batching, workers  # noqa: F401,E402
