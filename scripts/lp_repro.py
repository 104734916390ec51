import sys; sys.path.insert(0,'.')
import numpy as np, qpn_b200
from qpn_b200.polyhedra import LPSolver
eng = qpn_b200.Engine(0)
lp = LPSolver(eng)
rng = np.random.default_rng(0)
for d1, d2 in [(27, 20), (27, 30), (27, 40), (27, 43), (27, 50), (27, 60), (27, 69), (27, 70), (27, 80), (27,100), (40, 43), (10, 43), (27,44), (27,42)]:
    A = rng.normal(size=(d2, d1)); x0 = rng.normal(size=d1)
    l = A @ x0 - rng.uniform(0.1, 1, d2); u = A @ x0 + rng.uniform(0.1, 1, d2)
    try:
        r = lp.solve(rng.normal(size=d1), A, l, u)
        print(d1, d2, "n", d1 + 2 * d2, "status", r["status"], "big launches", eng.big_launches)
    except Exception as e:
        print(d1, d2, "n", d1 + 2 * d2, "ERROR", str(e)[:150])
