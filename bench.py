#!/usr/bin/env python
"""Benchmark of the QPNet equilibrium hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

One "step" = one pass of the hot path over one batch: `solve(qpn, inits)` for the
four-player Nash game (BASELINE.json configs[1]: examples/four_player_matrix_game.jl, 4,096
random initialisations per GPU).  Metric: equilibria/sec (whole job, all GPUs).

  value  device-timed (CUDA events on the launching stream), inputs already resident in HBM,
         L2 flushed between steps;
  e2e    the same work through the C-ABI call with pinned HOST buffers (H2D + kernel + D2H
         inside the timed region);
  roofline / cpu_baseline / clocks / gpu_launches as the bench contract asks.

Multi-GPU: independent instances are sharded over the ranks (weak scaling, 4,096 per GPU);
the only collective is the final NCCL all-gather of solutions / statuses / pivot counts.
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
WORKLOAD = "four_player_matrix_game Nash (edge_list=[]), random inits ~ U(-5,5)^8, one fused level-equilibrium launch per batch"
HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum of one level_equilibrium_kernel<32> launch at the default
# batch (4,096), from profiles/r1_final_level_kernel_ncu_full_summary.csv (ncu --set full, same command)
NCU_DRAM_BYTES_PER_LAUNCH_B4096 = 433_152 + 256
# same capture: l1tex__data_pipe_lsu_wavefronts_mem_shared.sum (each wavefront moves up to 128 B) and
# smsp__inst_executed.sum -- the two on-chip resources that actually bound the pivoting kernel
NCU_SMEM_WAVEFRONTS_PER_LAUNCH_B4096 = 3_884_408
NCU_WARP_INSTRUCTIONS_PER_LAUNCH_B4096 = 34_962_442
SM_COUNT = 148
# dram__bytes_read.sum + dram__bytes_write.sum of one level_equilibrium_big_kernel launch on the n = 256, m = 512
# monotone stress level at batch 148 (profiles/r1_big_level_kernel_ncu_full_summary.csv)
NCU_DRAM_BYTES_BIG_LEVEL_B148 = 259_560_047_000 + 250_557_664_000


def inits_for(rank, batch, step=0):
    rng = np.random.default_rng([0xB200, rank, step])
    return rng.uniform(-5.0, 5.0, (batch, 8))


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


def oracle_level(threads_hint=None):
    """The CPU restatement of the path (oracle/), used ONLY as the reported baseline."""
    from oracle import cport, examples as oex, qpn_ref
    net = oex.four_player_matrix_game()
    g, dec, par = qpn_ref.level_gavi(net, net.depth[1], {})
    import qpn_b200
    proj = qpn_b200.projection_vectors(qpn_b200.setup("four_player_matrix_game"))
    return cport.Level(8, [qpn_ref.node_view(net, p) for p in net.depth[1]], g, dec, par, 150, proj)


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(sample_batches=256, batch=4096):
    cores = cpu_cores()
    L = oracle_level()
    X = np.vstack([inits_for(10_000 + k, batch) for k in range(sample_batches)])
    L.solve(X[:batch], threads=cores)                       # warm
    t0 = time.perf_counter()
    r = L.solve(X, threads=cores)
    dt = time.perf_counter() - t0
    assert r["solved"].all()
    return {"value": len(X) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(X)} four_player equilibria ({sample_batches} batches of {batch}), C oracle, {cores} pthreads, {dt:.2f} s wall"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path.  Julia / PATH / OSQP are absent from this
    image, so the arm times the oracle port on all host cores."""
    if rank != 0:
        return
    cores = cpu_cores()
    L = oracle_level()
    B = args.batch
    for w in range(max(args.warmup, 1)):
        L.solve(inits_for(0, B, 1000 + w), threads=cores)
    times = []
    for k in range(args.steps):
        X = inits_for(0, B, k)
        t0 = time.perf_counter()
        r = L.solve(X, threads=cores)
        times.append(time.perf_counter() - t0)
        assert r["solved"].all()
    total = sum(times)
    val = B * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step": B,
                       "note": "CPU arm: one step = one batch of 4096 instances on the host cores (rank 0 only)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {B} equilibria, C oracle (oracle/qpn_oracle.c), {cores} pthreads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multilevel", action="store_true", help="skip the three-level robust_avoid extra (spawns host worker processes)")
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

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = qpn_b200.setup("four_player_matrix_game")
    solver = qpn_b200.BatchedSolver(net, device=local_rank)     # raises if libqpn_cuda / the GPU is missing
    eng = solver.engine
    level = solver.resident_level(1)
    B, nv, K, W = args.batch, net.n_vars, args.steps, args.warmup
    # An explicit stream: torch's default stream has handle 0, which the C ABI reads as "use the
    # handle's own stream" -- the events below must sit on the stream the kernel is launched on.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    # ---- device-resident arm -------------------------------------------------------------------
    x_in = [torch.from_numpy(inits_for(rank, B, k)).to(dev) for k in range(min(K, 8))]
    # Outputs of one rank live in ONE contiguous block [x_out | iters | pivots | solved] so that the
    # final gather of solutions / statuses / pivot counts is a single NCCL all-gather.
    o_it, o_pv, o_sol = B * nv * 8, B * nv * 8 + 4 * B, B * nv * 8 + 8 * B
    nbytes = (o_sol + B + 15) // 16 * 16
    block = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    x_out = block[:o_it].view(torch.float64).view(B, nv)
    iters = block[o_it:o_pv].view(torch.int32)
    pivots = block[o_pv:o_sol].view(torch.int32)
    solved = block[o_sol:o_sol + B]
    # Final gather of solutions / statuses / pivot counts.  Preferred form: the gathered buffer is
    # symmetric memory, every rank's level kernel stores its outputs straight into ITS slot of rank 0's buffer
    # (peer stores over NVLink from the kernel's own epilogue, no copy kernel), and a symmetric-memory barrier
    # on the same stream publishes them.  Falls back to one NCCL all-gather per step when symmetric memory
    # cannot be set up (or with QPN_BENCH_GATHER=nccl).
    gather_mode, hdl, out_base = "none", None, block.data_ptr()
    if world > 1:
        gather_mode = "nccl all_gather_into_tensor"
        if os.environ.get("QPN_BENCH_GATHER", "p2p") == "p2p":
            try:
                import torch.distributed._symmetric_memory as symm
                g_block = symm.empty(world * nbytes, dtype=torch.uint8, device=dev)
                hdl = symm.rendezvous(g_block, dist.group.WORLD)
                out_base = int(hdl.buffer_ptrs[0]) + rank * nbytes
                gather_mode = "kernel stores into rank 0's symmetric buffer over NVLink + symmetric-memory barrier"
            except Exception as e:                          # noqa: BLE001
                hdl = None
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({str(e)[:120]}); using NCCL all-gather", file=sys.stderr)
        if hdl is None:
            g_block = torch.empty(world * nbytes, dtype=torch.uint8, device=dev)
            out_base = block.data_ptr()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_dev(k):
        level.solve_dev(B, x_in[k % len(x_in)].data_ptr(), out_base, out_base + o_sol, out_base + o_it,
                        out_base + o_pv, None, stream.cuda_stream)

    def gather():
        if hdl is not None:
            hdl.barrier(channel=0)
        elif world > 1:
            dist.all_gather_into_tensor(g_block, block)

    for k in range(W):
        step_dev(k); gather()
    torch.cuda.synchronize()
    if hdl is not None:
        # rank 0 holds every rank's block: check them all there
        if rank == 0:
            gb = g_block.view(world, nbytes)
            assert bool(gb[:, o_sol:o_sol + B].bool().all()), "warm-up: not every instance (of every rank) reached an equilibrium"
        solved = g_block[o_sol:o_sol + B] if rank == 0 else None
        pivots = g_block[o_pv:o_sol].view(torch.int32) if rank == 0 else None
    else:
        assert bool(solved.bool().all()), "warm-up: not every instance reached an equilibrium"

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.zero_()                                   # evict the previous step's lines from L2 (not timed)
        ev[k][0].record(stream)
        step_dev(k)
        ev[k][1].record(stream)
        gather()
        ev[k][2].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - launches0
    t_kernel = sum(a.elapsed_time(b) for a, b, _ in ev) * 1e-3
    t_step = sum(a.elapsed_time(c) for a, _, c in ev) * 1e-3
    if hdl is not None and rank == 0:
        gb = g_block.view(world, nbytes)
        all_solved = bool(gb[:, o_sol:o_sol + B].bool().all())            # every rank's statuses, gathered on rank 0
        piv_host = gb[:, o_pv:o_sol].contiguous().view(torch.int32).cpu().numpy().ravel()
    elif hdl is not None:
        all_solved, piv_host = True, np.zeros(1)
    else:
        piv_host = pivots.cpu().numpy()
        all_solved = bool(solved.bool().all())

    # ---- end-to-end arm: pinned host buffers through the C ABI ---------------------------------
    hx = [torch.from_numpy(inits_for(rank, B, 100 + k)).pin_memory() for k in range(min(K, 8))]
    out = dict(x=torch.empty((B, nv), dtype=torch.float64).pin_memory().numpy(),
               solved=torch.empty(B, dtype=torch.uint8).pin_memory().numpy(),
               iters=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
               pivots=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), lam=None)
    hxn = [t.numpy() for t in hx]
    for k in range(W):
        level.solve(hxn[k % len(hxn)], out=out, want_lam=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_e2e = 0.0
    for k in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        level.solve(hxn[k % len(hxn)], out=out, want_lam=False)        # H2D + kernel + D2H + sync
        t_e2e += time.perf_counter() - t0
        assert out["solved"].all()
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([t_kernel, t_step, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_kernel, t_step, t_e2e = (float(v) for v in t.cpu())
        ok = torch.tensor([int(all_solved)], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        all_solved = bool(ok.item())

    # ---- secondary workload (reported under "extra", not the headline): the bottom level of
    # robust_avoid_simple (BASELINE.json configs[2]; LP-like nodes, lifted AVI n = 52, presolve on
    # every instance).  The full three-level solve is still a per-instance host recursion.
    extra = None
    try:
        ra = qpn_b200.setup("robust_avoid_simple")
        ra_solver = qpn_b200.BatchedSolver(ra, engine=eng)
        ra_level = ra_solver.resident_level(ra.num_levels())
        Br = 8192
        rng = np.random.default_rng([0xB200, rank, 7])
        Xr = np.tile(ra.default_initialization, (Br, 1))
        Xr[:, 0:6] += 0.5 * rng.normal(size=(Br, 6)); Xr[:, 6:12] = rng.uniform(-1, 1, (Br, 6))
        xr = torch.from_numpy(Xr).to(dev)
        xo = torch.empty_like(xr); so = torch.empty(Br, dtype=torch.uint8, device=dev)
        io = torch.empty(Br, dtype=torch.int32, device=dev); po = torch.empty(Br, dtype=torch.int32, device=dev)
        run = lambda: ra_level.solve_dev(Br, xr.data_ptr(), xo.data_ptr(), so.data_ptr(), io.data_ptr(), po.data_ptr(), None, stream.cuda_stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tr = 0.0
        for _ in range(5):
            flush.zero_(); e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize(); tr += e0.elapsed_time(e1) * 1e-3
        extra = {"robust_avoid_bottom_level": {"value": 5 * Br / tr, "unit": UNIT, "batch": Br, "ms_per_launch": 1e3 * tr / 5,
                                               "all_solved": bool(so.bool().all()), "p50_pivots_per_solve": float(np.median(po.cpu().numpy())),
                                               "note": "per GPU, device-timed; level 3 of 3 only"}}
    except Exception as e:                                      # the headline must not depend on the extra
        extra = {"robust_avoid_bottom_level": {"error": str(e)[:200]}}

    # ---- BASELINE.json configs[4]: the n = 256, m = 512 monotone stress QP as a one-node QPNet (lifted level AVI
    # n = 1,536) on the global-memory tableau path: one wave of persistent CTAs (one instance per SM).
    try:
        ms = qpn_b200.setup("monotone_stress")
        ms_solver = qpn_b200.BatchedSolver(ms, engine=eng)
        ms_level = ms_solver.resident_level(1)
        info = ms_level.info()
        Bm = SM_COUNT
        rng = np.random.default_rng([0xB200, rank, 9])
        xm = torch.from_numpy(ms.default_initialization + rng.normal(size=(Bm, ms.n_vars))).to(dev)
        xo = torch.empty_like(xm); so = torch.empty(Bm, dtype=torch.uint8, device=dev)
        io = torch.empty(Bm, dtype=torch.int32, device=dev); po = torch.empty(Bm, dtype=torch.int32, device=dev)
        run = lambda: ms_level.solve_dev(Bm, xm.data_ptr(), xo.data_ptr(), so.data_ptr(), io.data_ptr(), po.data_ptr(), None, stream.cuda_stream)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
        tm = e0.elapsed_time(e1) * 1e-3
        pv = po.cpu().numpy().astype(np.float64)
        # algorithmic HBM bytes: every pivot an instance runs itself reads and writes the live tableau once
        # (16 B per entry, n rows x live columns); the plan's pivots were run once for the whole batch
        own_pivots = float((pv - info["plan_pivots"]).sum())
        alg = own_pivots * 16.0 * info["n"] * info["ncol0"]
        peak_m, _ = measured_hbm_peak()
        extra["monotone_stress_n256_m512"] = {
            "value": Bm / tm, "unit": UNIT, "batch": Bm, "ms_per_launch": 1e3 * tm, "all_solved": bool(so.bool().all()),
            "p50_pivots_per_solve": float(np.median(pv)), "lifted_n": info["n"], "live_columns": info["ncol0"], "plan_pivots": info["plan_pivots"],
            "path": "global-memory tableau" if info["big"] else "shared-memory tableau",
            "roofline": {"bound": "hbm", "achieved": NCU_DRAM_BYTES_BIG_LEVEL_B148 / tm / 1e9, "peak": peak_m, "unit": "GB/s",
                         "frac": NCU_DRAM_BYTES_BIG_LEVEL_B148 / tm / 1e9 / peak_m, "traffic": NCU_DRAM_BYTES_BIG_LEVEL_B148,
                         "dense_upper_bound_bytes": alg,
                         "note": "achieved = DRAM bytes of this launch measured by ncu (same batch, same data) / event-timed duration; the dense "
                                 "bound 16 B x rows x live columns x pivots overstates it by 12x: the rows of free basics (1,024 of 1,536) are frozen "
                                 "after the plan and never swept, pivots are queued four deep and swept in one pass, rows with a zero entering "
                                 "entry and column pairs with zero pivot-row entries are skipped, and L2 serves 75 % of the sector requests; "
                                 "the sweep is load-latency bound per SM (L1 36 %, issue 27 %, DRAM 31 % of the copy peak), not HBM bound"},
            "note": "per GPU, device-timed; verify -> solve_qep -> verify fused in level_equilibrium_big_kernel"}
        ms_solver.close()
    except Exception as e:
        extra["monotone_stress_n256_m512"] = {"error": str(e)[:200]}

    # ---- BASELINE.json configs[3]: the synthetic 3-level chain (n = 64 per node).  Its levels are measured one at a
    # time -- here the bottom level (64 own variables, 136 parameters, lifted level AVI n = 256 of which 64 rows are
    # swept): the full three-level solve needs solution graphs of 64-variable nodes, which neither the host mirror nor
    # (per its README) the reference produces in usable time.
    try:
        ch = qpn_b200.setup("synthetic_chain")
        ch_solver = qpn_b200.BatchedSolver(ch, engine=eng)
        ch_level = ch_solver.resident_level(ch.num_levels())
        cinfo = ch_level.info()
        Bc = 4096
        rng = np.random.default_rng([0xB200, rank, 13])
        xc = torch.from_numpy(ch.default_initialization + 0.7 * rng.normal(size=(Bc, ch.n_vars))).to(dev)
        xo = torch.empty_like(xc); so = torch.empty(Bc, dtype=torch.uint8, device=dev)
        io = torch.empty(Bc, dtype=torch.int32, device=dev); po = torch.empty(Bc, dtype=torch.int32, device=dev)
        run = lambda: ch_level.solve_dev(Bc, xc.data_ptr(), xo.data_ptr(), so.data_ptr(), io.data_ptr(), po.data_ptr(), None, stream.cuda_stream)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize()
        tc = e0.elapsed_time(e1) * 1e-3
        extra["synthetic_chain_bottom_level"] = {
            "value": Bc / tc, "unit": UNIT, "batch": Bc, "ms_per_launch": 1e3 * tc, "all_solved": bool(so.bool().all()),
            "p50_pivots_per_solve": float(np.median(po.cpu().numpy())), "lifted_n": cinfo["n"], "live_columns": cinfo["ncol0"],
            "plan_pivots": cinfo["plan_pivots"],
            "path": "global-memory engine" if cinfo["big"] else "shared-memory tableau (swept rows only: 64 of 256)",
            "note": "per GPU, device-timed; level 3 of 3 only"}
        ch_solver.close()
    except Exception as e:                                      # noqa: BLE001
        extra["synthetic_chain_bottom_level"] = {"error": str(e)[:200]}

    # ---- BASELINE.json configs[2] in full: three-level robust_avoid_simple solves (vertex exploration on), the host
    # recursion of solve_base! sharded over worker processes that are all served by this rank's engine handle
    # (workers.py).  Host-bound (piece generation / set operations in the Python mirror): reported as an extra.
    if world == 1 and not args.no_multilevel:
        try:
            ra3 = qpn_b200.setup("robust_avoid_simple", seed=3)
            nw = max(1, min(cpu_cores() - 1, 15))
            Bf = 128 * nw
            rng = np.random.default_rng([0xB200, rank, 11])
            Xf = np.tile(ra3.default_initialization, (Bf, 1))
            Xf[:, 0:6] += 0.3 * rng.normal(size=(Bf, 6)); Xf[:, 6:12] = rng.uniform(-1, 1, (Bf, 6))
            with qpn_b200.MultilevelPool(ra3, nw, engine=eng) as pool:
                pool.solve(Xf[:4 * nw])                                   # processes up, piece memos warm
                st = {}
                l0 = eng.launches
                t0 = time.perf_counter(); res = pool.solve(Xf, stats=st); tf = time.perf_counter() - t0
            extra["robust_avoid_three_levels"] = {
                "value": Bf / tf, "unit": UNIT, "batch": Bf, "host_workers": nw, "wall_s": tf,
                "solved_fraction": float(np.mean([r["solved"] for r in res])), "device_launches": int(eng.launches - l0),
                "device_requests_before_regrouping": int(st.get("requests", 0)),
                "note": "per GPU, host clock, inputs and results in host memory; every numeric step on the device, the recursion of "
                        "solve_base! in worker processes (host-bound)"}
        except Exception as e:                                  # noqa: BLE001
            extra["robust_avoid_three_levels"] = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        # algorithmic HBM bytes of one launch: x in, x/solved/iters/pivots out, + the level's matrices once
        alg_bytes = B * (8 * nv + 8 * nv + 1 + 4 + 4) + 8 * (32 * 32 + 3 * 32 + 4 * (2 * 8 + 2 + 2 * 8 + 4)) + 8 * 4 * nv
        achieved = alg_bytes / (t_kernel / K) / 1e9
        line = {
            "metric": METRIC, "value": world * B * K / t_step, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_step / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "n_vars": nv, "avi_size": 32, "l2": "flushed between steps (256 MB write)",
                       "timing": "CUDA events on the launching stream, max over ranks", "all_solved": all_solved, "gather": gather_mode,
                       "p50_pivots_per_solve": float(np.median(piv_host)), "kernel_ms_per_step": 1e3 * t_kernel / K},
            "e2e": {"value": world * B * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": B * nv * 8, "d2h_bytes_per_step": B * (nv * 8 + 1 + 4 + 4),
                    "timing": "host clock around the synchronous C-ABI call (qpn_level_equilibrium_resident), pinned buffers"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH_B4096 if B == 4096 else None,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel": "level_equilibrium_kernel<32>", "peak_source": peak_src,
                         "note": "the fused pivoting kernel is issue/latency bound in shared memory (ncu: IPC 1.7/SM, fp64 pipe 10 %, "
                                 "0 % tensor), not HBM bound; see DESIGN.md 5 and profiles/"},
            "clocks": clocks,
            "extra": extra,
        }
        if B == 4096:
            # On-chip view (not the contract's HBM/tensor roofline, which this latency/issue-bound fp64 kernel
            # cannot approach): shared-memory bandwidth 128 B/clk/SM and issue rate 4 warp-instr/clk/SM at the
            # SM clock seen during the run; counts per launch from the committed ncu capture.
            clk = (clocks or {}).get("sm_mhz") or 1965.0
            t_k = t_kernel / K
            smem_peak = SM_COUNT * 128 * clk * 1e6 / 1e12
            issue_peak = SM_COUNT * 4 * clk * 1e6 / 1e12
            line["roofline_onchip"] = {
                "smem": {"achieved": NCU_SMEM_WAVEFRONTS_PER_LAUNCH_B4096 * 128 / t_k / 1e12, "peak": smem_peak, "unit": "TB/s",
                         "frac": NCU_SMEM_WAVEFRONTS_PER_LAUNCH_B4096 * 128 / t_k / 1e12 / smem_peak},
                "issue": {"achieved": NCU_WARP_INSTRUCTIONS_PER_LAUNCH_B4096 / t_k / 1e12, "peak": issue_peak, "unit": "T warp-instr/s",
                          "frac": NCU_WARP_INSTRUCTIONS_PER_LAUNCH_B4096 / t_k / 1e12 / issue_peak},
                "source": "profiles/r1_final_level_kernel_ncu_full_summary.csv"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
