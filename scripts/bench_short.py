"""dev: device-timed ms per solve (solve only, stage buckets) for A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
width = int(sys.argv[2]) if len(sys.argv) > 2 else 8
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, 2)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0), ls_width=width)
b.enable_timing(True)
ts = []
for _ in range(6):
    b.set_option(m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0)); b.initialization(); b.solve(); b.synchronize()
    ts.append(b.timing())
t = ts[-1]
print("ls_width", width, "solve ms (last 4):", [round(x["total_ms"], 2) for x in ts[2:]], {k: round(t[k], 2) for k in ("rollout_ms", "linearize_ms", "bws_ms", "linesearch_ms", "other_ms")}, flush=True)
