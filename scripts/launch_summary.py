"""dev: profiles/rN_launches.csv (ncu --metrics gpu__time_duration.sum --csv) -> profiles/rN_launches_summary.md

    python scripts/launch_summary.py [round=2] [bench record, default profiles/rN_bench_end_of_round.json]
"""
import csv, collections, json, os, sys
RND = sys.argv[1] if len(sys.argv) > 1 else "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(l for l in open(os.path.join(ROOT, "profiles", f"r{RND}_launches.csv")) if not l.startswith("==")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: t = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError: continue
    t *= {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "ns": 1e-6, "us": 1e-3, "ms": 1}.get(r[ix["Metric Unit"]], 1e-6)
    name = r[ix["Kernel Name"]].split("(")[0].replace("mhpc::<unnamed>::", "").replace("unnamed>::", "").replace("mhpc::", "").replace("void ", "")
    if name == "k_lin": name = f"k_lin grid {r[ix['Grid Size']]}"
    agg[name] += t; cnt[name] += 1
tot = sum(agg.values())
L = [f"# ncu launch list, round {RND} (end): `ncu --metrics gpu__time_duration.sum --clock-control none -c 640 --csv python bench.py --steps 1 --warmup 1 --cpu-sample 8`", "",
     f"{sum(cnt.values())} launches captured (the first ~6 solves of the run), total {tot:.1f} ms of kernel time (cold-cache, serialised: compare shares).", "",
     "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in agg.most_common(20):
    L.append(f"| `{k[:60]}` | {cnt[k]} | {v:.2f} | {100*v/tot:.1f} % |")
BENCH = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", f"r{RND}_bench_end_of_round.json")
d = json.load(open(BENCH))
st = d["stage_ms"]; t = st["total_ms"]
L += ["", "`k_lin grid (32, 79, 1)` (two kernels) = back/front-stance Jacobian knots `k_lin<BS>`, `k_lin<FS>`; `(32, 198, 1)` = flight knots `k_lin<FL>`; `(32, 1, 1)` (two kernels) = impact Jacobians at the two touchdowns (side stream in a real run); `(32, 362, 1)` = light knots (rigid-body + stance-terminal). `k_rollout<1, 4>` = 4-trial launches (line-search chunks and the forward sweeps, which run as four identical trials), `k_rollout<1, 2>` = the 2-trial remainder of the line search; all staged through shared memory. `k_copy_nominal` runs on a side stream underneath `k_lin`/`k_bws` in a real run. `k_bwsT<0>` = the first launch of a sweep stage (all instances, first rungs of the ladder in the kernel), `k_bwsT<1>` = the task launches of the ladder rounds.",
      f"Shares agree with the CUDA-event stage buckets bench.py reports (`stage_ms` of `profiles/r{RND}_bench_end_of_round.json`, {t:.1f} ms per solve: backward sweep {100*st['bws_ms']/t:.0f} %, line search {100*st['linesearch_ms']/t:.0f} %, linearise {100*st['linearize_ms']/t:.0f} %, forward sweeps {100*st['rollout_ms']/t:.0f} %, rest {100*st['other_ms']/t:.0f} %)."]
open(os.path.join(ROOT, "profiles", f"r{RND}_launches_summary.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L[4:14]))
