"""Flat-array model format: a QPNet as JSON, so that nets built by the reference's symbolic front
end (`add_constraint!` / `add_qp!` / `add_edges!`, /root/reference/src/programs.jl:147-285) can feed this
engine without Symbolics on the engine side.  The Julia hook that writes the same
schema from a `QPNet` is `julia/QPNCuda.jl: export_qpnet`.

Schema (all indices 1-based, as Julia stores them; matrices as CSC triplets `I, J, V` + shape; an infinite
bound is `null`):

  {"format": "qpn-b200/1", "n_vars": n, "variables": [names...],
   "qps": {"<id>": {"Q": {"m":n,"n":n,"I":[],"J":[],"V":[]}, "q": [...], "k": 0.0,
                    "constraint_indices": [...], "var_indices": [...]}},
   "constraints": {"<id>": {"A": {csc}, "l": [...], "u": [...], "rl": [0/1 strict...], "ru": [...],
                            "group_mapping": {"<player>": group}}},
   "edges": [[parent, child], ...],            # the minimal adjacency (network_edges)
   "options": {QPNetOptions fields}, "default_initialization": [...]}
"""
import json

import numpy as np

from .model import INF, Poly, QP, QPNet, QPNetOptions

FORMAT = "qpn-b200/1"


def _csc(M):
    M = np.asarray(M, dtype=float)
    J, I = np.nonzero(M.T)                      # column-major order, as SparseMatrixCSC iterates
    return {"m": int(M.shape[0]), "n": int(M.shape[1]), "I": (I + 1).tolist(), "J": (J + 1).tolist(), "V": M[I, J].tolist()}


def _dense(c):
    M = np.zeros((c["m"], c["n"]))
    if c["I"]:
        M[np.asarray(c["I"]) - 1, np.asarray(c["J"]) - 1] = c["V"]
    return M


def _bounds_out(v):
    return [None if np.isinf(x) else float(x) for x in v]


def _bounds_in(v, sign):
    return np.array([sign * INF if x is None else x for x in v], dtype=float)


def net_to_dict(net):
    opts = {k: getattr(net.options, k) for k in QPNetOptions.__dataclass_fields__}
    if opts["levels_to_remove_subsets"] is not None:
        opts["levels_to_remove_subsets"] = sorted(opts["levels_to_remove_subsets"])
    return {
        "format": FORMAT, "n_vars": net.n_vars, "variables": list(net.names),
        "qps": {str(i): {"Q": _csc(qp.Q), "q": qp.q.tolist(), "k": float(qp.k),
                         "constraint_indices": list(qp.constraint_indices), "var_indices": [v + 1 for v in qp.var_indices]}
                for i, qp in net.qps.items()},
        "constraints": {str(i): {"A": _csc(P.A), "l": _bounds_out(P.l), "u": _bounds_out(P.u), "rl": P.rl.astype(int).tolist(),
                                 "ru": P.ru.astype(int).tolist(), "group_mapping": {str(k): v for k, v in getattr(net, "group_map", {}).get(i, {}).items()}}
                        for i, P in net.constraints.items()},
        "edges": [[int(i), int(j)] for i, js in sorted(net.network_edges.items()) for j in js],
        "options": opts, "default_initialization": np.asarray(net.default_initialization, float).tolist(),
    }


def export_net(net, path):
    with open(path, "w") as f:
        json.dump(net_to_dict(net), f)


def net_from_dict(d):
    if d.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} model: format = {d.get('format')!r}")
    n = int(d["n_vars"])
    net = QPNet(("x", n))
    if d.get("variables"):
        net.names = list(d["variables"])
    for cid, c in sorted(d["constraints"].items(), key=lambda kv: int(kv[0])):
        net.constraints[int(cid)] = Poly(_dense(c["A"]), _bounds_in(c["l"], -1), _bounds_in(c["u"], +1),
                                         np.asarray(c.get("rl", [0] * len(c["l"])), bool), np.asarray(c.get("ru", [0] * len(c["u"])), bool))
    for pid, q in sorted(d["qps"].items(), key=lambda kv: int(kv[0])):
        net.qps[int(pid)] = QP(_dense(q["Q"]), np.asarray(q["q"], float), float(q.get("k", 0.0)), [int(c) for c in q["constraint_indices"]],
                               [int(v) - 1 for v in q["var_indices"]])
    net.add_edges([tuple(e) for e in d.get("edges", [])])
    net.assign_constraint_groups({int(c): {int(k): v for k, v in cc.get("group_mapping", {}).items()} for c, cc in d["constraints"].items()})
    opts = dict(d.get("options", {}))
    if opts.get("levels_to_remove_subsets") is not None:
        opts["levels_to_remove_subsets"] = set(opts["levels_to_remove_subsets"])
    opts.pop("shared_variable_mode", None) if not isinstance(opts.get("shared_variable_mode"), str) else None
    net.set_options(**opts)
    net.default_initialization = np.asarray(d.get("default_initialization", np.zeros(n)), float)
    return net


def load_net(path):
    with open(path) as f:
        return net_from_dict(json.load(f))
