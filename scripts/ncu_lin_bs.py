"""Launch only k_lin<BS> a few times at the bench batch size (for the ncu --set full capture of the dominant kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, 2)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0))
b.initialization(); b.forward_sweep_dynamics_only(0.0); b.synchronize()
print("k_lin<BS> ms", b.time_kernel(10, 2), "k_lin<FL> ms", b.time_kernel(12, 2), flush=True)
