#!/usr/bin/env python
"""bench.py - batched MHPC solves/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--instances 4096]

Workload (config.workload): 4096 independent MHPC instances per GPU, default gait and 4 WB + 4 RB phase split,
seeded perturbed initial states (SURVEY.md 8(d) config-2 generator, seed 20260101), a FIXED HSDDP budget of
max_AL_iter x max_DDP_iter = 2 x 3 inner iterations (DDP_thresh = AL_thresh = -1: never converge early).
A "step" = MHPCLocomotion::initialization() + solve_mhpc() of the whole batch.  All arithmetic FP64.

value : solves/s with the problems resident in HBM (device-timed, CUDA events, max over ranks)
e2e   : solves/s through the host-facing call mhpc_batch_solve_host: problems from pinned host memory (H2D), solve,
        results + the reference's execution bag back to the host (D2H: x,u,y,K,du of the first two whole-body phases, 180 knots per
        instance, --exec-knots), all inside the timed region;  e2e_first10: the same with the 10 knots one control period uses
roofline     : the dominant kernel k_bwsT (backward sweep on the FP64 tensor cores) - algorithmic bytes per launch / launch time,
               timed live, against the measured HBM copy bandwidth (MEASURED_PEAKS.json); traffic from profiles/r2_bwsT_ncu_traffic.json
roofline_lin_*: the linearisation kernels: EXECUTED flops / DFMA peak measured live (the reference formulation's flop count over the
               same time is kept under reference_formulation_*)
hsddp_iteration_ms(_max), iteration_ms: CUDA-event time per HSDDP iteration of the batch (the 10 ms control period of the north star)
strong_scaling: BASELINE config 5, one 32768-instance batch over 1/2/4/8 GPUs (--strong-total; recorded in profiles/r2_strong_scaling.json)
cpu_baseline : the reference's own sources (oracle/_ref; else the oracle port) on one host core, bounded sample
--impl reference : the same on all host cores (one process per core), bounded sample per step
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "batched MHPC solves/sec (4096 inst, fixed HSDDP iters)"
UNIT = "solves/s"
# algorithmic flops of one default-horizon linearisation (158 stance + 198 flight WB knots, 356 RB, 2 impacts)
LIN_FLOP_PER_INSTANCE = 6896848.0
BS_FLOP_PER_KNOT = 27569.0
# bytes the backward sweep must move per instance (records in + K,du,G out), default horizon: 356 WB step records (175 dbl),
# 4 WB terminal (148), 352 RB step (20), 4 RB terminal (6) read; 712 x (56+4+14) written
BWS_BYTES_PER_INSTANCE = 8.0 * (356 * 175 + 4 * 148 + 352 * 20 + 4 * 6 + 712 * 74)
OPT_KW = dict(DDP_thresh=-1.0, AL_thresh=-1.0)


def workload_name(n):
    return (f"{n} instances/GPU, default bounding gait, 4 WB + 4 RB phases (720 knots), perturbed x0 (seed 20260101), "
            "fixed 2x3 HSDDP iterations; step = initialization + solve")


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def _cpu_kind():
    """'reference' = the reference's own solver sources (oracle/_ref/libmhpc_reference_fast.so, compiled against the Eigen
    stand-in); 'port' = the oracle restatement, used only when that library is not available."""
    from oracle import reference_py
    return "reference" if reference_py.available("fast") else "port"


def _oracle_solve_chunk(args):
    specs_idx, variant = args
    kind = _cpu_kind()
    t = time.perf_counter()
    if kind == "reference":
        from oracle.reference_py import ReferenceSolver
        for i in specs_idx:
            s = ReferenceSolver(_spec_of(i), variant, **OPT_KW)
            s.initialization()
            s.solve()
    else:
        from helpers import to_orc_problem
        from oracle.oracle_py import OracleSolver
        for i in specs_idx:
            s = OracleSolver(to_orc_problem(_spec_of(i), **OPT_KW), variant=variant)
            s.initialization()
            s.solve()
    return time.perf_counter() - t


def _spec_of(i):
    from helpers import perturbed_x0
    return dict(n_wbphase=4, n_fbphase=4, dt_wb=0.001, dt_fb=0.001, cmode=1, vel=1.5, height=0.0,
                gait_mode=[1, 2, 3, 4], gait_timing=[0.08, 0.1, 0.08, 0.1], x0=perturbed_x0(i))


def cpu_baseline_single(sample):
    """oracle, one core, first `sample` instances of the workload."""
    t = _oracle_solve_chunk((list(range(sample)), "fast"))
    return sample / t


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    kind = _cpu_kind()
    if kind == "reference":
        from oracle import reference_py
        reference_py.lib("fast")
    else:
        from oracle.oracle_py import build
        build("fast")
    cores = os.cpu_count() or 1
    per_core = max(1, args.ref_sample // cores)
    sample = per_core * cores
    chunks = [(list(range(c * per_core, (c + 1) * per_core)), "fast") for c in range(cores)]
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_oracle_solve_chunk, chunks)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.instances) + f" [CPU arm: {sample} of these instances per step]"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sample} instances per step ({per_core} per core, one process per core); "
                                       + ("the reference's own solver sources + generated C, -O3 -mavx2 -mfma, built against the Eigen stand-in oracle/eigen_shim"
                                          if kind == "reference" else "oracle restatement, -O3 -mavx2 -mfma")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--instances", type=int, default=4096, help="instances per GPU")
    # execution bag of the e2e call: the reference copies the first two whole-body phases (MHPCLocomotion.cpp:178-194),
    # N0 + N1 = 80 + 100 knots of the default gait; --exec-knots 10 gives the lighter figure reported as e2e_first10
    ap.add_argument("--exec-knots", type=int, default=180, help="knots of the execution bag returned by the e2e call")
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--ref-sample", type=int, default=256)
    # line-search trials evaluated side by side per launch.  At 4096 instances the accepted step size is among the first
    # three trials in 5 of the 6 iterations (scripts/nbws_hist.py) and eight trial rollouts per instance load the FP64 pipe
    # (2 warps/scheduler): four slots per launch, with a second launch for the few instances that need trials 5+, is faster
    # (line search 8.5 ms instead of 12.0 ms per solve); the outcome is the same (tests: line-search width invariance)
    ap.add_argument("--ls-width", type=int, default=4)
    # BASELINE config 5: ONE batch of this many instances split over the ranks (strong scaling); default: weak scaling,
    # --instances per GPU
    ap.add_argument("--strong-total", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import mhpc_minimal_env_b200 as m
    from helpers import make_specs, to_gpu_problem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION/INFO write it to stdout) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        import torch.distributed as dist
        saved_stdout = os.dup(1)          # ... and anything else the communicator set-up prints goes to stderr as well
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()                # creates the communicator now, not inside the timed region
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from mhpc_minimal_env_b200.sharding import gather_result_table, shard_range
    n = args.instances if not args.strong_total else (args.strong_total + world - 1) // world
    # weak scaling: n instances per rank; the job of n*world instances is sharded by instance into contiguous blocks
    # (no collective inside the solve)
    g_begin, g_end = shard_range(n * world, rank, world)
    assert g_end - g_begin == n
    specs = [_spec_of(i) for i in range(g_begin, g_end)]
    problems = [to_gpu_problem(s) for s in specs]
    opt = m.default_option(**OPT_KW)
    batch = m.BatchedMHPC(problems, opt, device=local, ls_width=args.ls_width)
    stream = torch.cuda.Stream()          # a real (non-NULL) stream shared by the library and the CUDA events below
    torch.cuda.set_stream(stream)
    batch.set_stream(stream.cuda_stream)

    def step():
        batch.set_option(opt)      # a fresh MHPCLocomotion per solve: solve() leaves _option modified (MultiPhaseDDP.cpp:178-183,273-277)
        batch.initialization()
        batch.solve()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # final result gather (the only communication of the job): per-rank result table -> every rank, over NCCL
    res_arr = batch.results_array()
    npad = (n + 31) // 32 * 32
    table = torch.zeros(6 * npad, device="cuda", dtype=torch.float64)
    batch.results_to_device(table.data_ptr(), table.numel())
    torch.cuda.synchronize()
    full = gather_result_table(table.view(6, npad)[:, :n].contiguous(), n * world, world, dist)
    n_ok = int((full[3] == 2).sum().item())

    # ---- end-to-end through the host-facing call, host buffers, copies inside the timed region ----
    import ctypes as C
    from mhpc_minimal_env_b200.capi import Problem, Result
    L = m.load()
    dp = C.POINTER(C.c_double)
    pin_p = torch.empty(n * C.sizeof(Problem), dtype=torch.uint8).pin_memory()
    C.memmove(pin_p.data_ptr(), C.addressof(batch._problems), n * C.sizeof(Problem))
    pin_r = torch.empty(n * C.sizeof(Result), dtype=torch.uint8).pin_memory()

    def measure_e2e(ek):
        pin_s = torch.empty(n * ek * 22, dtype=torch.float64).pin_memory()
        pin_g = torch.empty(n * ek * 60, dtype=torch.float64).pin_memory()

        def e2e_step():
            rc = L.mhpc_batch_solve_host(batch.h, C.cast(pin_p.data_ptr(), C.POINTER(Problem)), C.cast(pin_r.data_ptr(), C.POINTER(Result)),
                                         2, C.cast(pin_s.data_ptr(), dp), C.cast(pin_g.data_ptr(), dp), ek)
            if rc < 0:
                raise RuntimeError(L.mhpc_last_error().decode())

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        return n * world / float(t_e2e.item()), n * C.sizeof(Result) + n * ek * (22 + 60) * 8

    ek = args.exec_knots
    e2e_value, d2h = measure_e2e(ek)
    e2e10_value, d2h10 = measure_e2e(10) if ek != 10 else (e2e_value, d2h)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = n * C.sizeof(Problem)

    line = None
    if rank == 0:
        # ---- per-kernel-family timing and rooflines (rank 0 only, outside the timed region) ----
        batch.enable_timing(True)
        step()
        batch.synchronize()
        tm = batch.timing()
        itm = batch.iteration_timing()
        batch.enable_timing(False)
        lin_ms = batch.time_kernel(0, 5)
        lin_bs_ms = batch.time_kernel(10, 5)      # the single dominant kernel: k_lin<BS> (back-stance Jacobian knots)
        batch.backward_sweep(1.0)                 # large regularisation: every instance completes the sweep
        bws_ms = batch.time_kernel(1, 3)
        fp64_peak = m.measure_fp64_peak(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # flops the device forms actually execute (the Jacobians are re-derived by implicit differentiation,
        # codegen/implicit.py: ~5x fewer operations than the reference's flat CasADi DAGs that SURVEY.md 8(d) counts)
        exe = {}
        try:
            exe = json.load(open(os.path.join(ROOT, "mhpc_minimal_env_b200", "csrc", "gen", "gen_meta.json")))["_regalloc"]["device_flops"]
        except Exception:
            pass
        n_bs_knots = 79                                   # back-stance step knots per default horizon (phase 0)
        wb_cost_flop = 150.0                              # running-cost / barrier partials of a WB step knot
        exe_bs = (exe.get("Dyn_BS_par", 4880) + wb_cost_flop) * n_bs_knots * n
        exe_lin = n * (79 * (exe.get("Dyn_BS_par", 4880) + wb_cost_flop) + 79 * (exe.get("Dyn_FS_par", 4880) + wb_cost_flop)
                       + 198 * (exe.get("Dyn_FL_par", 3710) + wb_cost_flop) + exe.get("Imp_F_par", 15166) + exe.get("Imp_B_par", 15166)
                       + 352 * (exe.get("FBDynamics_par", 27) + 60.0))
        lin_tflops = LIN_FLOP_PER_INSTANCE * n / (lin_ms * 1e-3) / 1e12
        bws_gbs = BWS_BYTES_PER_INSTANCE * n / (bws_ms * 1e-3) / 1e9
        bs_flop = BS_FLOP_PER_KNOT * n_bs_knots * n       # 27 569 flop per WB-stance knot (reference formulation)
        bs_tflops = bs_flop / (lin_bs_ms * 1e-3) / 1e12

        def ncu_traffic(name):   # dram bytes of ONE launch at this batch size from the committed ncu --set full capture
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", name)))
                if prof.get("instances") == n:
                    return prof["dram_bytes_per_launch"]
            except Exception:
                pass
            return None
        # the dominant kernel of a step is the backward sweep (stage_ms below): k_bwsT, one warp per instance on the FP64 tensor
        # cores; algorithmic bytes = knot records in + K, du, G out
        roofline = {"kernel": "k_bwsT (impact-aware Riccati sweep over 712 knots on the FP64 tensor cores, one attempt, all instances)", "bound": "hbm",
                    "achieved": bws_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": bws_gbs / hbm_peak,
                    "traffic": ncu_traffic("r2_bwsT_ncu_traffic.json"), "algorithmic_bytes": BWS_BYTES_PER_INSTANCE * n,
                    "ms_per_launch": bws_ms, "us_per_knot": 1e3 * bws_ms / 712,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "note": "712 dependent knots per instance; ncu (profiles/r2_ncu_summary.md): shared-memory data pipe 51 %, issue slots 45 %, "
                            "FP64 tensor pipe 37 % + FP64 pipe 21 % of peak (stalls: fixed-latency dependencies 34 %, math-pipe throttle 24 %); DMMA and "
                            "DFMA have the same FP64 peak on B200 (profiles/r2_dmma_microbench.txt)"}
        # FP64 kernels: frac = flops the device form EXECUTES / measured DFMA peak.  (The Jacobians are re-derived by implicit
        # differentiation, ~5x fewer operations than the reference's flat CasADi DAGs; the reference formulation's flop count of
        # SURVEY.md 8(d) over the same time is kept under reference_formulation_* for comparison with the CPU path only.)
        roofline_lin_bs = {"kernel": "k_lin<BS> (Dyn_BS_par + Euler A,B + cost/barrier partials, 79 back-stance knots x instances)",
                           "bound": "fp64", "achieved": exe_bs / (lin_bs_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                           "frac": exe_bs / (lin_bs_ms * 1e-3) / 1e12 / fp64_peak,
                           "reference_formulation_tflops": bs_tflops, "reference_formulation_frac": bs_tflops / fp64_peak,
                           "traffic": ncu_traffic("r2_lin_bs_ncu_traffic.json"), "algorithmic_bytes": 1352.0 * n_bs_knots * n,
                           "hbm_frac_algorithmic": 1352.0 * n_bs_knots * n / (lin_bs_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_launch": lin_bs_ms,
                           "peak_source": "FP64 FMA probe measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)"}
        roofline_lin_group = {"kernel": "linearisation launch group (k_lin<BS,FS,FL,IMPF,IMPB,light>)", "bound": "fp64",
                              "achieved": exe_lin / (lin_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": exe_lin / (lin_ms * 1e-3) / 1e12 / fp64_peak,
                              "reference_formulation_tflops": lin_tflops, "reference_formulation_frac": lin_tflops / fp64_peak,
                              "ms_per_launch": lin_ms, "us_per_horizon_linearize": 1e3 * lin_ms / n}
        strong = None
        try:
            strong = json.load(open(os.path.join(ROOT, "profiles", "r2_strong_scaling.json")))
        except Exception:
            pass
        cpu_v = cpu_baseline_single(args.cpu_sample)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong_total else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(n), "instances_per_gpu": n, "hsddp_budget": "2x3 (DDP_thresh=AL_thresh=-1)",
                           "ls_width": args.ls_width, "l2": "working set (>10 GB/GPU) far exceeds the 126 MB L2; no explicit flush",
                           "sharding": "by instance, contiguous blocks per rank, one final NCCL all_gather of the result table"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": f"mhpc_batch_solve_host: pinned host problems in; results + the execution bag of the reference (x,u,y,K,du of the first two "
                                f"whole-body phases, {ek} knots per instance, MHPCLocomotion.cpp:178-194) out"},
                "e2e_first10": {"value": e2e10_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h10,
                                "note": "the same call returning only the first 10 knots of the execution bag (what one 10 ms control period consumes)"},
                "gpu_launches": int(tm["n_launches"] + 3) * args.steps,
                "clocks": sampler.summary(),
                "roofline": roofline, "roofline_lin_bs": roofline_lin_bs, "roofline_lin_group": roofline_lin_group,
                "stage_ms": {k: tm[k] for k in ("total_ms", "rollout_ms", "linearize_ms", "bws_ms", "linesearch_ms", "other_ms")},
                # the north star's unit of work: ONE HSDDP iteration over the horizon for the whole batch (10 ms control period)
                "hsddp_iteration_ms": tm["total_ms"] / 6.0,
                "hsddp_iteration_ms_max": max((r["total_ms"] for r in itm), default=None),
                "iteration_ms": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()} for r in itm],
                "strong_scaling": strong,
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": _cpu_kind(),
                                 "sample": f"first {args.cpu_sample} instances of the workload, reference sources (-O3 -mavx2 -mfma, Eigen stand-in), one core; "
                                           f"host has {os.cpu_count()} cores -> x{os.cpu_count()} = {cpu_v * (os.cpu_count() or 1):.1f} solves/s if it scaled perfectly"},
                "solves_finished": n_ok, "device_bytes": batch.device_bytes()}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
