import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mhpc_minimal_env_b200 as m
from oracle.oracle_py import OracleSolver
from helpers import *
np.set_printoptions(precision=6, linewidth=200)
specs = make_specs(4, 2)
gpu = m.BatchedMHPC([to_gpu_problem(s) for s in specs])
gpu.initialization()
for i in range(4):
    o = OracleSolver(to_orc_problem(specs[i])); o.initialization()
    for nm in ("X_NOM","U_NOM","Y_NOM"):
        g = gpu.field(i, 0, nm); r = o.field(0, nm)
        print(i, nm, relerr(g, r))
        if relerr(g,r) > 1e-9:
            print(" gpu k0", g[0], "\n orc k0", r[0], "\n gpu k1", g[1], "\n orc k1", r[1], "\n gpu k40", g[40], "\n orc k40", r[40])
