"""dev: per-launch durations of one solve of the bench workload from an ncu launch list (csv) -> grouped, in launch order.

    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file X.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0
    python scripts/launch_list.py X.csv [first_launch] [n_launches]
"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: t = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError: continue
    t *= {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "ns": 1e-6, "us": 1e-3, "ms": 1}.get(r[ix["Metric Unit"]], 1e-6)
    name = r[ix["Kernel Name"]].split("(")[0].replace("mhpc::<unnamed>::", "").replace("mhpc::", "").replace("void ", "")
    out.append((name, r[ix["Grid Size"]], t))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = lo + (int(sys.argv[3]) if len(sys.argv) > 3 else len(out))
agg = collections.Counter(); cnt = collections.Counter()
for k, (n, g, t) in enumerate(out[lo:hi]):
    print(f"{lo + k:5d} {n:28s} {g:18s} {t:9.4f} ms")
    agg[n] += t; cnt[n] += 1
print("---- totals")
for n, t in agg.most_common():
    print(f"{n:28s} x{cnt[n]:4d} {t:9.3f} ms")
print("sum", round(sum(agg.values()), 3))
