import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(64, 2)], m.default_option())
for name, fn in [("init", b.initialization), ("fs", lambda: b.forward_sweep(0.0)), ("upd", b.update_nominal_trajectory), ("bws", lambda: b.backward_sweep(1.0)), ("fit", b.forward_iteration)]:
    try:
        fn(); b.synchronize(); print(name, "ok", flush=True)
    except Exception as e:
        print(name, "FAILED", e, flush=True); break
