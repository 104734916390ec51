"""Test helper: an object with the Engine methods the host logic uses, backed by the CPU oracle.
It lets the `-m "not gpu"` suite exercise the multi-level driver, the polyhedral operations and
the solution-graph code; the `-m gpu` suite runs the same logic on the real engine."""
import numpy as np

from oracle import cport


class OracleEngine:
    launches = 0

    def gavi_solve(self, g, w, z0, presolve=True, max_pivots=0):
        g = g if isinstance(g, dict) else g.source
        z0 = np.atleast_2d(np.asarray(z0, dtype=float))
        B = z0.shape[0]
        w = np.asarray(w, dtype=float).reshape(B, -1)
        outs = [cport.gavi_solve(g, z0[b], w[b], presolve=presolve, max_pivots=max_pivots) for b in range(B)]
        return dict(z=np.array([o["z"] for o in outs]), z_full=np.array([o["z_full"] for o in outs]),
                    status=np.array([o["status"] for o in outs], np.int32), pivots=np.array([o["pivots"] for o in outs], np.int32),
                    basis=np.array([o["basis"] for o in outs]))

    def verify_solution(self, node, x, tol=1e-4):
        x = np.atleast_2d(x)
        outs = [cport.verify_solution(*node, x[b], tol) for b in range(len(x))]
        return (np.array([o[0] for o in outs]), np.array([o[1] for o in outs]).reshape(len(x), -1),
                np.array([o[2] for o in outs], np.int32), np.array([o[3] for o in outs]).reshape(len(x), -1))

    def comp_indices(self, g, z, w, tol=1e-2):
        z = np.atleast_2d(z); w = np.asarray(w, dtype=float).reshape(len(z), -1)
        return np.array([cport.comp_indices(g, z[b], w[b], tol) for b in range(len(z))])

    def halfspace_in(self, polys, x, tol=1e-6):
        x = np.atleast_2d(x)
        return np.array([[cport.halfspace_in(P[0], P[1], P[2], x[j], tol) if len(P[1]) else True for P in polys] for j in range(len(x))])
