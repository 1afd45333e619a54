"""dev: time ONE full successful backward sweep (large regularisation so that no instance fails)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, cfg)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0))
b.initialization(); b.forward_sweep(0.0); b.update_nominal_trajectory()
ok = b.backward_sweep(1.0)
print("n", n, "cfg", cfg, "success", int(ok.sum()), "full bws ms", round(b.time_kernel(1, 3), 3), "lin ms", round(b.time_kernel(0, 3), 3), "rollout ms", round(b.time_kernel(2, 3), 3), flush=True)
