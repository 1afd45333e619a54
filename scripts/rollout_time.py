"""dev: time one line-search rollout launch: python scripts/rollout_time.py [instances=4096] [trial slots per launch=8]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, 2)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0), ls_width=W)
b.initialization(); b.solve(); b.synchronize()
print(f"rollout{W} ms", round(b.time_kernel(2, 3), 3), flush=True)
