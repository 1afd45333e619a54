"""dev: aggregate an `ncu --page source --csv --print-source sass,cuda` export by source line range (regions given as
file:lo-hi=name ...); prints instruction share, sample share and warp instructions per (instance, knot)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])          # instances x knots
regs = []
for a in sys.argv[3:]:
    loc, name = a.split("=")
    f, rng = loc.split(":")
    lo, hi = rng.split("-")
    regs.append((f, int(lo), int(hi), name))
cur = None; ix = None
inst = collections.Counter(); samp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        ix = {}
        for i, h in enumerate(r): ix.setdefault(h, i)
        continue
    if ix is None or not r[0].isdigit(): continue
    ln = int(r[0])
    def num(k):
        try: return int(r[ix[k]])
        except Exception: return 0
    name = cur
    for f, lo, hi, nm in regs:
        if f == cur and lo <= ln <= hi: name = nm; break
    inst[name] += num('Instructions Executed'); samp[name] += num('# Samples')
ti = sum(inst.values()); ts = sum(samp.values())
print("total warp instructions", ti, "per unit %.1f" % (ti / units))
for k, v in inst.most_common():
    print(f"{k:36s} inst {100*v/ti:5.1f}%  samples {100*samp[k]/max(ts,1):5.1f}%   per unit {v/units:6.1f}")
