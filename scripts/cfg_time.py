"""dev: solve time of one BASELINE config (fixed 2x3 budget): python scripts/cfg_time.py [cfg=4] [instances=4096] [ls_width=8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
opt = m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0)
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, cfg)], opt, ls_width=W)
b.enable_timing(True)
for _ in range(3):
    b.set_option(opt)      # a fresh MHPCLocomotion per solve, as in bench.py
    b.initialization(); b.solve(); b.synchronize()
tm = b.timing()
print("cfg", cfg, "n", n, "W", W, {k: round(tm[k], 2) for k in ("total_ms", "bws_ms", "linesearch_ms", "linearize_ms", "rollout_ms")}, flush=True)
