"""dev: A/B of the two lane mappings of the backward sweep (MHPC_BWS16=0/1): one full sweep timed, K/du/G compared bit for
bit, then a whole fixed-budget solve (traces must be identical)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mhpc_minimal_env_b200 as m
from helpers import make_specs, to_gpu_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
probs = [to_gpu_problem(s) for s in make_specs(n, cfg)]
out = {}
kinds = (sys.argv[3].split(",") if len(sys.argv) > 3 else ["8", "T"])
for v in kinds:
    os.environ["MHPC_BWS"] = v
    b = m.BatchedMHPC(probs, m.default_option(DDP_thresh=-1.0, AL_thresh=-1.0), ls_width=4)
    b.initialization(); b.forward_sweep(0.0); b.update_nominal_trajectory()
    ok = b.backward_sweep(1.0)
    t = b.time_kernel(1, 5)
    pick = [0, 1, 2, 3, n // 2, n - 1]
    f = {(i, p, nm): b.field(i, p, nm) for i in pick for p in range(len(b.schedule(i)[0])) for nm in ("K", "DU", "G")}
    sc = b.scalars()["exp_cost_change"].copy()
    b.enable_timing(True)
    b.initialization(); b.solve(); b.synchronize()
    tm = b.timing()
    b.initialization(); b.solve(); b.synchronize()
    tm = b.timing()
    recs = [b.records(i) for i in range(0, n, max(1, n // 256))]
    out[v] = (f, sc, recs, b.results_array().tobytes())
    print("BWS=" + v, "success", int(ok.sum()), "full sweep ms %.3f" % t, "solve: total %.2f bws %.2f ls %.2f lin %.2f" % (tm["total_ms"], tm["bws_ms"], tm["linesearch_ms"], tm["linearize_ms"]), flush=True)
    b.close()
A, B = kinds[0], kinds[1]
f0, f1 = out[A][0], out[B][0]
nd = sum(0 if np.array_equal(f0[k], f1[k]) else 1 for k in f0)
worst = max(float(np.max(np.abs(f0[k] - f1[k]) / (np.max(np.abs(f0[k])) + 1e-300))) for k in f0)
print("fields differing bitwise:", nd, "of", len(f0), "worst rel diff %.2e" % worst)
ints = lambda rr: [[(r["iter_AL"], r["iter_DDP"], r["n_bws"], r["n_fit"], r["reg"], r["eps"]) for r in x] for x in rr]
dv0, dv1 = out[A][1], out[B][1]
print("dV worst rel diff %.2e" % float(np.max(np.abs(dv0 - dv1) / np.abs(dv0))), "integer traces identical:", ints(out[A][2]) == ints(out[B][2]),
      "worst rel cost diff %.2e" % max(abs(a["cost"] - b["cost"]) / abs(a["cost"]) for x, y in zip(out[A][2], out[B][2]) for a, b in zip(x, y)))
for k in f0:
    e = float(np.max(np.abs(f0[k] - f1[k]) / (np.max(np.abs(f0[k])) + 1e-300)))
    if e > 1e-9: print("  field", k, "rel diff %.2e" % e)
