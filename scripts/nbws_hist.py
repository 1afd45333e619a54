"""dev: distribution of backward-sweep attempts (n_bws) per DDP iteration over the bench workload."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, 2)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0))
b.initialization(); b.solve(); b.synchronize()
hist = collections.defaultdict(collections.Counter)
hfit = collections.defaultdict(collections.Counter)
for i in range(n):
    for k, r in enumerate(b.records(i)):
        hist[k][r["n_bws"]] += 1
        hfit[k][r["n_fit"]] += 1
for k in sorted(hist):
    print("iteration", k, "n_bws", dict(sorted(hist[k].items())), "n_fit", dict(sorted(hfit[k].items())), flush=True)
