"""Evidence for the regularisation-retry tail of the backward sweep (MultiPhaseDDP.cpp:209-227).

For each BASELINE workload and each inner (DDP) iteration: the histogram of attempts per instance (n_bws), and per
attempt the backward step at which the failed attempts stopped (as a fraction of the instance's horizon) - i.e. how
much of a full sweep every failed attempt costs.  Writes profiles/r2_nbws_hist.json.

    python scripts/nbws_hist.py [n_instances] [out.json]
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_nbws_hist.json")
out = {"n_instances": n, "option": "fixed 2x3 iterations (DDP_thresh = AL_thresh = -1)", "workloads": {}}
for name, cfg, lsw in (("config2_bench", 2, 4), ("config3_all_wb", 3, 8), ("config4_varied", 4, 8)):
    b = m.BatchedMHPC([to_gpu_problem(s) for s in make_specs(n, cfg)], m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0), ls_width=lsw)
    b.enable_bws_trace(True)
    b.initialization()
    b.solve()
    b.synchronize()
    tr = b.bws_trace()                       # [n, rec, att]
    nk = np.array([sum(b.schedule(i)[1]) for i in range(n)]) if cfg == 4 else np.full(n, sum(b.schedule(0)[1]))
    recs = [b.records(i) for i in range(n)]
    nrec = max(len(r) for r in recs)
    w = {"iterations": []}
    for k in range(nrec):
        have = np.array([len(r) > k for r in recs])
        nb = np.array([r[k]["n_bws"] if len(r) > k else 0 for r in recs])
        nf = np.array([r[k]["n_fit"] if len(r) > k else 0 for r in recs])
        it = {"record": k, "iter_AL": recs[int(np.argmax(have))][k]["iter_AL"], "iter_DDP": recs[int(np.argmax(have))][k]["iter_DDP"],
              "instances": int(have.sum()),
              "n_bws_hist": {int(a): int(c) for a, c in sorted(collections.Counter(nb[have].tolist()).items())},
              "n_fit_hist": {int(a): int(c) for a, c in sorted(collections.Counter(nf[have].tolist()).items())},
              "attempts": []}
        work = np.zeros(n)
        for a in range(tr.shape[2]):
            t = tr[:, k, a]
            made = have & (t != 0)
            if not made.any():
                break
            failed = made & (t > 0)
            frac = np.where(failed, t / np.maximum(nk, 1), 1.0)     # fraction of the horizon walked by this attempt
            work += np.where(made, frac, 0.0)
            q = np.quantile(frac[failed], [0, 0.25, 0.5, 0.75, 1.0]).round(3).tolist() if failed.any() else None
            it["attempts"].append({"attempt": a + 1, "made": int(made.sum()), "failed": int(failed.sum()),
                                   "fail_depth_quantiles_min_q25_med_q75_max": q,
                                   "fail_step_hist_by_100": {int(x): int(c) for x, c in sorted(collections.Counter(((t[failed] - 1) // 100 * 100).tolist()).items())}})
        it["sweep_equivalents_mean"] = float(work[have].mean())
        it["sweep_equivalents_max"] = float(work[have].max())
        # what the one-launch retry loop costs: a warp (4 consecutive slots) walks max over its instances per attempt
        it["n_bws_max"] = int(nb.max())
        w["iterations"].append(it)
        print(name, "record", k, "n_bws", it["n_bws_hist"], "mean work", round(it["sweep_equivalents_mean"], 2), "max", round(it["sweep_equivalents_max"], 2), flush=True)
        for a in it["attempts"]:
            print("    attempt", a["attempt"], "made", a["made"], "failed", a["failed"], "depth", a["fail_depth_quantiles_min_q25_med_q75_max"], flush=True)
    out["workloads"][name] = w
    b.close()
with open(out_path, "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote", out_path)
