import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
for cfg, n in ((2, 5), (4, 7), (3, 3)):
    b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, cfg)])
    b.initialization(); b.solve(); b.synchronize()
    r, st, gn = b.solve_host(exec_knots=4)
    print(cfg, [round(x.cost, 3) for x in r])
    b.close()
