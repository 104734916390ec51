#!/usr/bin/env python
"""Benchmark of the QPNet equilibrium hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): examples/robust_avoid_simple.jl in
full -- three levels (ego -> adversaries -> separating planes), exploration_vertices = 10 -- for a batch of 65,536
perturbed instances, sharded over the GPUs (strong scaling: 65,536 / N per GPU).  One "step" = one
`solve(qpn, inits::Matrix)` = one `qpn_net_solve_batched` per rank: the whole recursion of solve_base! for every
instance, every numeric step on the device.  Metric: equilibria/sec (whole job, all GPUs).

  value  inputs (inits) and outputs (x) resident in HBM (qpn_net_solve_batched_dev), timed with CUDA events,
         L2 flushed between steps, max over ranks, the one result exchange per solve included;
  e2e    the same through the host-pointer C-ABI call (pinned host buffers; H2D + every launch + D2H inside the
         timed region), host clock; e2e_pageable: the same with pageable buffers (a Julia Matrix{Float64});
  roofline / cpu_baseline / clocks / gpu_launches as the bench contract asks; `extra` carries the other BASELINE
  configs (four_player_matrix_game batch of 4,096; bottom levels; the n = 256 / m = 512 stress shape).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "QPNet equilibria/sec (batched AVI solves)"
UNIT = "equilibria/s"
WORKLOAD = ("robust_avoid_simple in full (examples/robust_avoid_simple.jl defaults: 2 obstacles, 5 faces, 3 levels, "
            "exploration_vertices=10, num_projections=5; stand-in problem data seed 3), batch of perturbed instances "
            "(xe, xo ~ default + N(0, 0.5^2), ue, uo ~ U(-1,1)), one solve(qpn, inits) = solve_base! on every level per instance")
L2_NOTE = "GPU arm: flushed between steps (256 MB write); CPU arm: not applicable"
DATA_SEED = 3
HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
SM_COUNT = 148
# ncu --set full captures of the two dominant kernels of the headline (profiles/r2_*): DRAM bytes (read + write),
# shared-memory wavefronts and warp instructions per unit at the launch sizes named there
NCU = {}
try:
    with open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")) as _f:
        NCU = json.load(_f)
except Exception:       # noqa: BLE001
    NCU = {}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- the CPU restatement of the path (oracle/), used ONLY as the reported baseline / reference arm -------------------
def oracle_net_binding(net, threads):
    """The native state machine of csrc/net/ built against the C oracle (oracle/net_oracle.cpp): the same host logic
    as the product, every numeric step on the host cores, `threads` pthreads, none of the product's kernels."""
    import ctypes as C
    from qpn_b200 import netsolve
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libqpn_net_oracle.so"))
    return netsolve.NetBinding(net, lib, prefix="qpo_net_", threads=threads)


def cpu_baseline(sample=2, batch=65536):        # ~ 28 s of CPU work on the 16 host threads of the B200 box
    import qpn_b200
    cores = cpu_cores()
    net = qpn_b200.setup("robust_avoid_simple", seed=DATA_SEED)
    nb = oracle_net_binding(net, cores)
    nb.solve_arrays(qpn_b200.examples.robust_avoid_batch(net, 8192, seed=999))          # pieces memoised, threads up
    Xs = [qpn_b200.examples.robust_avoid_batch(net, batch, seed=2000 + k) for k in range(sample)]
    t0 = time.perf_counter()
    solved = 0
    for X in Xs:
        solved += int(nb.solve_arrays(X)["solved"].sum())
    dt = time.perf_counter() - t0
    n = sample * batch
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} full three-level robust_avoid equilibria ({sample} batches of {batch}), same host state machine as the GPU arm "
                      f"(cohorts, shared memo of pieces) with the C oracle as numeric backend, {cores} pthreads, {dt:.2f} s wall, "
                      f"solved fraction {solved / n:.4f}"}


def cpu_no_memo_sample(instances=24):
    """The reference's own structure -- one solve per instance, nothing shared between instances (every solve rebuilds
    its pieces: local_piece / project / LPs) -- on the C oracle, one instance per host thread."""
    import concurrent.futures as cf
    import qpn_b200
    cores = cpu_cores()
    net = qpn_b200.setup("robust_avoid_simple", seed=DATA_SEED)
    X = qpn_b200.examples.robust_avoid_batch(net, instances, seed=777)

    def one(b):
        nb = oracle_net_binding(net, 1)
        r = nb.solve_arrays(X[b:b + 1])
        nb.close()
        return bool(r["solved"][0])
    oracle_net_binding(net, 1).close()                                                   # library built and loaded
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=cores) as ex:
        ok = list(ex.map(one, range(instances)))
    dt = time.perf_counter() - t0
    return {"value": instances / dt, "unit": UNIT, "cores": cores, "instances": instances, "wall_s": dt, "solved_fraction": float(np.mean(ok)),
            "note": "per-instance solves with a fresh net object each (no memo shared between instances), as the reference's solve() works; "
                    "C oracle numerics, one instance per host thread"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path.  Julia / PATH / OSQP are absent from this image (SURVEY.md F2,
    F4), so the arm times the port: the same host logic on the C oracle, all host cores, a persistent thread pool per
    call of the whole batch (threads live for the 65,536 instances of a step)."""
    if rank != 0:
        return
    import qpn_b200
    cores = cpu_cores()
    B = args.batch
    net = qpn_b200.setup("robust_avoid_simple", seed=DATA_SEED)
    nb = oracle_net_binding(net, cores)
    nb.solve_arrays(qpn_b200.examples.robust_avoid_batch(net, min(B, 8192), seed=999))
    for w in range(max(args.warmup, 1)):
        nb.solve_arrays(qpn_b200.examples.robust_avoid_batch(net, B, seed=1000 + w))
    Xs = [qpn_b200.examples.robust_avoid_batch(net, B, seed=k) for k in range(min(args.steps, 4))]
    times, solved = [], 0
    for k in range(args.steps):
        t0 = time.perf_counter()
        r = nb.solve_arrays(Xs[k % len(Xs)])
        times.append(time.perf_counter() - t0)
        solved += int(r["solved"].sum())
    total = sum(times)
    val = B * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": B, "l2": L2_NOTE},
            "detail": {"solved_fraction": solved / (B * args.steps),
                       "note": "CPU arm: one step = the whole batch on the host cores (rank 0 only); same native host state machine as the GPU "
                               "arm (csrc/net/) with oracle/qpn_oracle.c as numeric backend"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {B} full three-level robust_avoid equilibria, {cores} pthreads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- extras: the other BASELINE configs, measured on this rank's GPU ----------------------------------------------------
def extras(qpn_b200, torch, eng, dev, stream, flush, rank):
    extra = {}
    ev = lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def timed(run, reps):
        run(); torch.cuda.synchronize()
        t = 0.0
        for _ in range(reps):
            e0, e1 = ev()
            flush.zero_(); e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
            t += e0.elapsed_time(e1) * 1e-3
        return t / reps

    def level_run(net, level, X):
        solver = qpn_b200.BatchedSolver(net, engine=eng)
        lv = solver.resident_level(level)
        B = len(X)
        xd = torch.from_numpy(X).to(dev)
        xo = torch.empty_like(xd); so = torch.empty(B, dtype=torch.uint8, device=dev)
        io = torch.empty(B, dtype=torch.int32, device=dev); po = torch.empty(B, dtype=torch.int32, device=dev)
        run = lambda: lv.solve_dev(B, xd.data_ptr(), xo.data_ptr(), so.data_ptr(), io.data_ptr(), po.data_ptr(), None, stream.cuda_stream)
        return solver, lv, run, so, po

    # BASELINE configs[3] END TO END at the node size whose solution graphs are tractable: the synthetic 3-level chain with 6
    # variables per node (26 variables), 4,096 instances through qpn_net_solve_batched (host buffers, host clock)
    try:
        from qpn_b200.netsolve import NetBinding
        ch = qpn_b200.setup("synthetic_chain", n=6, levels=3, n_params=8)
        B = 4096
        Xc = np.tile(ch.default_initialization, (B, 1)) + np.random.default_rng([0xB200, rank, 3]).normal(size=(B, ch.n_vars))
        nbc = NetBinding(ch, eng.lib, "qpn_net_", handle=eng.h, threads=2)
        t0 = time.perf_counter(); nbc.solve_arrays(Xc); cold = time.perf_counter() - t0
        nbc.solve_arrays(Xc)
        t0 = time.perf_counter()
        for _ in range(3):
            rc = nbc.solve_arrays(Xc)
        tw = (time.perf_counter() - t0) / 3
        extra["synthetic_chain_end_to_end_n6"] = {"value": B / tw, "unit": UNIT, "batch": B, "ms_per_solve": 1e3 * tw, "cold_first_solve_s": cold,
                                                  "solved_fraction": float(rc["solved"].mean()),
                                                  "mean_passes_per_level": rc["level_iters"].mean(0).round(2).tolist(),
                                                  "note": "three levels x 6 variables per node + 8 parameters, solve(qpn, inits) end to end (the whole "
                                                          "recursion, solution graphs included); the first solve of a net also builds and memoises its "
                                                          "geometry (cold_first_solve_s); n = 64 per node is out of reach for the solution graphs"}
        nbc.close()
    except Exception as e:                                  # noqa: BLE001
        extra["synthetic_chain_end_to_end_n6"] = {"error": str(e)[:200]}

    # examples/four_player_matrix_game.jl with edges (the example sweeps the power set of edge lists): the four-level chain
    # 1 -> 2 -> 3 -> 4, 4,096 random initialisations through qpn_net_solve_batched (host buffers, host clock)
    try:
        from qpn_b200.netsolve import NetBinding
        fh = qpn_b200.setup("four_player_matrix_game", edge_list=[(1, 2), (2, 3), (3, 4)])
        B = 4096
        Xh = np.random.default_rng([0xB200, rank, 4]).uniform(-5.0, 5.0, (B, 8))
        nbh = NetBinding(fh, eng.lib, "qpn_net_", handle=eng.h, threads=2)
        t0 = time.perf_counter(); nbh.solve_arrays(Xh); cold = time.perf_counter() - t0
        nbh.solve_arrays(Xh)
        t0 = time.perf_counter()
        for _ in range(5):
            rh = nbh.solve_arrays(Xh)
        th = (time.perf_counter() - t0) / 5
        extra["four_player_four_level_chain_b4096"] = {"value": B / th, "unit": UNIT, "batch": B, "ms_per_solve": 1e3 * th, "cold_first_solve_s": cold,
                                                      "solved_fraction": float(rh["solved"].mean()),
                                                      "mean_passes_per_level": rh["level_iters"].mean(0).round(2).tolist(),
                                                      "note": "edge_list = [(1,2),(2,3),(3,4)]: four levels, solution graphs at three of them; "
                                                              "solve(qpn, inits) end to end"}
        nbh.close()
    except Exception as e:                                  # noqa: BLE001
        extra["four_player_four_level_chain_b4096"] = {"error": str(e)[:200]}

    # BASELINE configs[1]: four_player_matrix_game, 4,096 random initialisations, one fused level launch
    try:
        fp = qpn_b200.setup("four_player_matrix_game")
        B = 4096
        X = np.random.default_rng([0xB200, rank, 0]).uniform(-5.0, 5.0, (B, 8))
        solver, lv, run, so, po = level_run(fp, 1, X)
        t = timed(run, 10)
        hx = torch.from_numpy(X).pin_memory().numpy()
        out = dict(x=torch.empty((B, 8), dtype=torch.float64).pin_memory().numpy(), solved=torch.empty(B, dtype=torch.uint8).pin_memory().numpy(),
                   iters=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), pivots=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), lam=None)
        for _ in range(3):
            lv.solve(hx, out=out, want_lam=False)
        t0 = time.perf_counter()
        for _ in range(10):
            lv.solve(hx, out=out, want_lam=False)
        te = (time.perf_counter() - t0) / 10
        extra["four_player_matrix_game_b4096"] = {"value": B / t, "unit": UNIT, "batch": B, "ms_per_launch": 1e3 * t, "e2e_value": B / te,
                                                  "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                                  "note": "BASELINE configs[1]; per GPU, device-timed (e2e_value: pinned host buffers through the C ABI)"}
        solver.close()
    except Exception as e:                                      # noqa: BLE001 -- the headline must not depend on an extra
        extra["four_player_matrix_game_b4096"] = {"error": str(e)[:200]}
    # bottom level of the headline network alone (level 3 of 3), fused level kernel
    try:
        ra = qpn_b200.setup("robust_avoid_simple", seed=DATA_SEED)
        X = qpn_b200.examples.robust_avoid_batch(ra, 8192, seed=7)
        solver, lv, run, so, po = level_run(ra, ra.num_levels(), X)
        t = timed(run, 5)
        extra["robust_avoid_bottom_level"] = {"value": len(X) / t, "unit": UNIT, "batch": len(X), "ms_per_launch": 1e3 * t,
                                               "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                               "note": "per GPU, device-timed; level 3 of 3 only (fused verify -> solve_qep -> verify kernel)"}
        solver.close()
    except Exception as e:                                      # noqa: BLE001
        extra["robust_avoid_bottom_level"] = {"error": str(e)[:200]}
    # BASELINE configs[3]: synthetic 3-level chain, n = 64 per node -- its bottom level (lifted n = 256, 64 rows swept)
    try:
        ch = qpn_b200.setup("synthetic_chain")
        X = ch.default_initialization + 0.7 * np.random.default_rng([0xB200, rank, 13]).normal(size=(4096, ch.n_vars))
        solver, lv, run, so, po = level_run(ch, ch.num_levels(), X)
        info = lv.info()
        t = timed(run, 2)
        extra["synthetic_chain_bottom_level"] = {"value": len(X) / t, "unit": UNIT, "batch": len(X), "ms_per_launch": 1e3 * t,
                                                  "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                                  "lifted_n": info["n"], "live_columns": info["ncol0"], "plan_pivots": info["plan_pivots"],
                                                  "note": "per GPU, device-timed; level 3 of 3 only"}
        solver.close()
    except Exception as e:                                      # noqa: BLE001
        extra["synthetic_chain_bottom_level"] = {"error": str(e)[:200]}
    # BASELINE configs[4]: n = 256, m = 512 monotone stress QP as a one-node QPNet (lifted level AVI n = 1,536)
    try:
        ms = qpn_b200.setup("monotone_stress")
        X = ms.default_initialization + np.random.default_rng([0xB200, rank, 9]).normal(size=(SM_COUNT, ms.n_vars))
        solver, lv, run, so, po = level_run(ms, 1, X)
        info = lv.info()
        t = timed(run, 1)
        extra["monotone_stress_n256_m512"] = {"value": len(X) / t, "unit": UNIT, "batch": len(X), "ms_per_launch": 1e3 * t,
                                               "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                               "lifted_n": info["n"], "live_columns": info["ncol0"], "plan_pivots": info["plan_pivots"],
                                               "path": "global-memory tableau" if info["big"] else "shared-memory tableau",
                                               "note": "per GPU, device-timed; one wave of persistent CTAs"}
        solver.close()
        # ... and at the batch BASELINE.json names (4,096): 28 waves of the same persistent grid, one timed launch
        X4 = ms.default_initialization + np.random.default_rng([0xB200, rank, 10]).normal(size=(4096, ms.n_vars))
        solver, lv, run, so, po = level_run(ms, 1, X4)
        e0, e1 = ev()
        flush.zero_(); e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
        t4 = e0.elapsed_time(e1) * 1e-3
        extra["monotone_stress_n256_m512_b4096"] = {"value": len(X4) / t4, "unit": UNIT, "batch": len(X4), "s_per_launch": t4,
                                                     "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                                     "note": "BASELINE configs[4] at its own batch; per GPU, device-timed, one launch"}
        solver.close()
    except Exception as e:                                      # noqa: BLE001
        extra["monotone_stress_n256_m512"] = {"error": str(e)[:200]}
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="instances per step, whole job (sharded over the GPUs)")
    ap.add_argument("--threads", type=int, default=0, help="host threads per rank that drive the batch (0 = by core count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import qpn_b200
    from qpn_b200.netsolve import NetBinding
    from qpn_b200.sharding import BlockLayout, ResultGather, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = qpn_b200.setup("robust_avoid_simple", seed=DATA_SEED)
    eng = qpn_b200.Engine(local_rank)                           # raises if libqpn_cuda / the GPU is missing
    threads = args.threads or max(1, min(2, cpu_cores() // world))   # cohorts are split on the device: two streams hide the round trips
    nb = NetBinding(net, eng.lib, "qpn_net_", handle=eng.h, threads=threads)
    Btot, nv, nl, K, W = args.batch, net.n_vars, net.num_levels(), args.steps, args.warmup
    lo, hi = shard_range(Btot, rank, world)
    B = hi - lo
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    # ---- inputs: a few distinct batches, rotated over the steps; every rank generates the job's batch and keeps its shard
    nsets = min(max(K, W), 4)
    host_sets = [qpn_b200.examples.robust_avoid_batch(net, Btot, seed=k)[lo:hi].copy() for k in range(nsets)]
    dev_in = [torch.from_numpy(X).to(dev) for X in host_sets]
    dev_out = torch.empty((B, nv), dtype=torch.float64, device=dev)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    flags = dict(solved=pin((B,), torch.uint8), level_iters=pin((B, nl), torch.int32), error=pin((B,), torch.int32))
    # the one exchange per solve: x | solved | level_iters | error of every rank in one block
    gather = None
    if world > 1:
        layout = BlockLayout([("x", np.float64, (nv,)), ("solved", np.uint8, ()), ("level_iters", np.int32, (nl,)), ("error", np.int32, ())],
                             max(h - l for l, h in (shard_range(Btot, r, world) for r in range(world))))
        gather = ResultGather(layout, device=dev, group=None, mode=os.environ.get("QPN_BENCH_GATHER", "nccl"))
        xoff, soff = layout.fields[0][3], layout.fields[1][3]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_dev(k):
        nb.solve_dev(B, dev_in[k % nsets].data_ptr(), dev_out.data_ptr(), out=flags)
        if gather is not None:
            # x is already on the device; the flags are produced on the host: pack them behind x and exchange once
            blk = gather.block
            blk[xoff: xoff + B * nv * 8].view(torch.float64).copy_(dev_out.view(-1), non_blocking=True)
            hb = gather.host_block.numpy()
            layout.pack({"x": np.zeros((0, nv)), "solved": flags["solved"], "level_iters": flags["level_iters"], "error": flags["error"]}, hb)
            blk[soff:].copy_(gather.host_block[soff:], non_blocking=True)
            gather.exchange_device()
            stream.synchronize()

    for k in range(W):
        step_dev(k)
    torch.cuda.synchronize()
    solved_frac_warm = float(flags["solved"].mean())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    st0, pr0 = nb.stats(), nb.profile()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    solved_total = 0
    iters_sum = np.zeros(nl)
    for k in range(K):
        flush.zero_()                                   # evict the previous step's lines from L2 (not timed)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        step_dev(k)
        ev[k][1].record(stream)
        solved_total += int(flags["solved"].sum())
        iters_sum += flags["level_iters"].sum(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    st1, pr1 = nb.stats(), nb.profile()
    t_step = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    launches = st1["launches"] - st0["launches"]

    # ---- end-to-end arm: pinned host buffers through the host-pointer C-ABI call ------------------------------------
    hx = [torch.from_numpy(X).pin_memory().numpy() for X in host_sets]
    hout = dict(x=pin((B, nv), torch.float64), **flags)
    for k in range(2):
        nb.solve_arrays(hx[k % nsets], out=hout)
    if world > 1:
        dist.barrier()
    pe0 = nb.profile()
    t_e2e = 0.0
    for k in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = nb.solve_arrays(hx[k % nsets], out=hout)           # H2D + every launch of the solve + D2H + sync
        if gather is not None:
            gather.exchange({"x": r["x"], "solved": flags["solved"], "level_iters": flags["level_iters"], "error": flags["error"]})
        t_e2e += time.perf_counter() - t0
    pe1 = nb.profile()
    clocks = sampler.stop() if sampler else None
    # pageable host buffers (what a Julia Matrix{Float64} is): the same call, staged through the driver's bounce buffers
    t_page = 0.0
    reps_page = min(K, 3)
    for k in range(reps_page):
        Xp = np.array(host_sets[k % nsets])
        t0 = time.perf_counter()
        nb.solve_arrays(Xp)
        t_page += time.perf_counter() - t0

    # ---- per-kernel durations: one extra pass with every launch bracketed by CUDA events on its stream ----------------
    nb.set_option("profile", 1)
    pk0 = nb.profile()
    for k in range(2):
        nb.solve_dev(B, dev_in[k % nsets].data_ptr(), dev_out.data_ptr(), out=flags)
    pk1 = nb.profile()
    nb.set_option("profile", 0)
    kern = {k: {f: pk1[k][f] - pk0[k][f] for f in ("launches", "units", "ms")} for k in ("verify", "solve_qep", "member", "group", "cycle")}

    if world > 1:
        t = torch.tensor([t_step, t_e2e, t_page], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step, t_e2e, t_page = (float(v) for v in t.cpu())
        s = torch.tensor([solved_total, launches], dtype=torch.float64, device=dev)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        solved_total, launches = (float(v) for v in s.cpu())

    extra = None
    if rank == 0 and not args.no_extras:
        try:
            extra = extras(qpn_b200, torch, eng, dev, stream, flush, rank)
        except Exception as e:                                  # noqa: BLE001
            extra = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        nproj = net.options.num_projections
        dom = max(kern, key=lambda k: kern[k]["ms"])
        # algorithmic HBM bytes per unit of each kernel: slot index + x in (+ x out) + flags / masks / projections
        alg_unit = {"solve_qep": 4 + 8 * nv + 8 * nv + 5 + 8 * nproj + 8, "verify": 4 + 8 * nv + 1 + 16 + 8, "member": 4 + 8 * nv + 1 + 8,
                    "group": 4 + 4 + 2 * (8 + 4) + 4, "cycle": 4 + 8 * nproj * 2 + 4 + 1 + 8}
        kd = kern[dom]
        per_launch_units = kd["units"] / max(kd["launches"], 1)
        launch_s = kd["ms"] * 1e-3 / max(kd["launches"], 1)
        achieved = alg_unit[dom] * per_launch_units / launch_s / 1e9 if launch_s > 0 else 0.0
        total_kern_ms = sum(v["ms"] for v in kern.values())
        ncu_k = NCU.get(dom, {})
        line = {
            "metric": METRIC, "value": Btot * K / t_step, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_step / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": Btot, "l2": L2_NOTE},
            "detail": {"instances_per_gpu": B, "host_threads_per_rank": threads, "n_vars": nv, "levels": nl,
                       "timing": "CUDA events around the synchronous call on every rank, max over ranks; the one result exchange per solve included",
                       "solved_fraction": solved_total / (Btot * K), "mean_iterations_per_level": (iters_sum / (B * K)).tolist(),
                       "gather": (gather.mode + ": one block per rank per solve") if gather else "none (1 GPU)",
                       "rounds_per_step": (st1["rounds"] - st0["rounds"]) / K, "batched_calls_per_step": (st1["calls"] - st0["calls"]) / K,
                       "requests_per_step": (st1["requests"] - st0["requests"]) / K, "new_lps_in_timed_region": st1["lps"] - st0["lps"],
                       "host_ms_per_step_summed_over_threads": (st1["host_ns"] - st0["host_ns"]) / 1e6 / K,
                       "backend_ms_per_step_summed_over_threads": (st1["backend_ns"] - st0["backend_ns"]) / 1e6 / K,
                       "staged_h2d_bytes_per_step": (pr1["h2d_bytes"] - pr0["h2d_bytes"]) / K,
                       "staged_d2h_bytes_per_step": (pr1["d2h_bytes"] - pr0["d2h_bytes"]) / K},
            "e2e": {"value": Btot * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": (pe1["h2d_bytes"] - pe0["h2d_bytes"]) / K,
                    "d2h_bytes_per_step": (pe1["d2h_bytes"] - pe0["d2h_bytes"]) / K,
                    "timing": "host clock around the synchronous C-ABI call (qpn_net_solve_batched), pinned host buffers; bytes counted by the "
                              "library: inits in, round tables in, one representative's answers per cohort part out, x and outcome index out (this rank)"},
            "e2e_pageable": {"value": Btot * reps_page / t_page, "unit": UNIT, "steps": reps_page,
                             "note": "same call with pageable numpy buffers (a Julia Matrix{Float64} is pageable)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_k["dram_bytes_per_unit"] * per_launch_units) if ncu_k.get("dram_bytes_per_unit") else None,
                         "kernel": {"solve_qep": "net_qep_kernel", "verify": "net_verify_kernel", "member": "net_member_kernel",
                                    "group": "net_round_* + cub::DeviceSegmentedSort", "cycle": "net_cycle_kernel"}[dom],
                         "algorithmic_bytes_per_unit": alg_unit[dom], "units_per_launch": per_launch_units,
                         "avg_launch_ms": 1e3 * launch_s, "peak_source": peak_src,
                         "kernel_time_share": {k: (v["ms"] / total_kern_ms if total_kern_ms else 0.0) for k, v in kern.items()},
                         "kernel_ms_per_step": {k: v["ms"] / 2 for k, v in kern.items()},
                         "launches_per_step": {k: v["launches"] / 2 for k, v in kern.items()},
                         "note": "durations from a separate pass with every launch bracketed by CUDA events on its own stream (host threads run "
                                 "concurrent streams, so the per-kernel sums exceed the step time when launches overlap); the pivoting kernels are "
                                 "issue / shared-memory-latency bound (fp64 rank-1 updates in shared memory), not HBM bound: see roofline_onchip, "
                                 "DESIGN.md 5 and profiles/"},
            "clocks": clocks,
            "extra": extra,
        }
        if ncu_k.get("smem_wavefronts_per_unit"):
            clk = (clocks or {}).get("sm_mhz") or 1965.0
            smem_peak = SM_COUNT * 128 * clk * 1e6 / 1e12
            wf = ncu_k["smem_wavefronts_per_unit"] * per_launch_units
            line["roofline_onchip"] = {"smem": {"achieved": wf * 128 / launch_s / 1e12, "peak": smem_peak, "unit": "TB/s",
                                                "frac": wf * 128 / launch_s / 1e12 / smem_peak},
                                       "source": ncu_k.get("source")}
            if ncu_k.get("warp_instructions_per_unit"):
                issue_peak = SM_COUNT * 4 * clk * 1e6
                wi = ncu_k["warp_instructions_per_unit"] * per_launch_units
                line["roofline_onchip"]["issue"] = {"achieved": wi / launch_s / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-instructions/s",
                                                    "frac": wi / launch_s / issue_peak,
                                                    "note": "the average launch of a solve is smaller than the captured one: the tail rounds of a batch "
                                                            "hold a few cohorts and run at the latency of one dependent chain"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if extra is not None:
                try:
                    extra["cpu_reference_structure_no_shared_memo"] = cpu_no_memo_sample()
                except Exception as e:                          # noqa: BLE001
                    extra["cpu_reference_structure_no_shared_memo"] = {"error": str(e)[:200]}
        print(json.dumps(line), flush=True)
    nb.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
