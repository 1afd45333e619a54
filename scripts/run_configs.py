"""Time the BASELINE.json configs 2-5 on one GPU (device-timed, CUDA events inside the library): default thresholds
(the reference's own convergence flow) and the fixed 2x3 budget.  Writes gpurun_out/r2_configs.json (copied to profiles/ when kept)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem

rows = []
for name, cfg, n in (("config2 (1024, default split)", 2, 1024), ("config2 at 4096", 2, 4096), ("config3 (4096, all whole-body)", 3, 4096),
                     ("config4 (4096, varied splits/timings, 8-way line search)", 4, 4096), ("config2 at 16384", 2, 16384), ("config5 (32768 on ONE GPU)", 2, 32768)):
    t0 = time.time()
    probs = [to_gpu_problem(s) for s in make_specs(n, cfg)]
    for label, kw in (("default thresholds", {}), ("fixed 2x3", dict(DDP_thresh=-1.0, AL_thresh=-1.0))):
        opt = m.default_option(**kw)
        b = m.BatchedMHPC(probs, opt)
        b.enable_timing(True)
        for _ in range(2):
            b.set_option(opt)      # a fresh MHPCLocomotion per solve, as in bench.py: solve() leaves its option copy modified
            b.initialization(); b.solve(); b.synchronize()
        tm = b.timing()
        res = b.results_array()
        rows.append({"config": name, "options": label, "instances": n, "solve_ms": tm["total_ms"], "solves_per_s": n / (tm["total_ms"] * 1e-3),
                     "stage_ms": {k: round(tm[k], 3) for k in ("rollout_ms", "linearize_ms", "bws_ms", "linesearch_ms", "other_ms")},
                     "mean_inner_iterations": float(res["n_records"].mean()), "status_counts": {int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
                     "device_GB": b.device_bytes() / 1e9})
        print(rows[-1], flush=True)
        b.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r2_configs.json"), "w"), indent=1)
