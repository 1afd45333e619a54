"""dev: aggregate an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; ix = None
agg = collections.Counter(); inst = collections.Counter(); srcs = {}; stall = collections.defaultdict(collections.Counter)
KEYS = ('stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_mio', 'stall_branch_resolving', 'stall_math', 'stall_no_inst',
        'stall_not_selected', 'stall_selected', 'stall_barrier', 'stall_dispatch', 'stall_lg')
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        ix = {}
        for i, h in enumerate(r): ix.setdefault(h, i)
        continue
    if ix is None or not r[0].isdigit(): continue
    ln = int(r[0])
    def num(k):
        try: return int(r[ix[k]])
        except Exception: return 0
    agg[(cur, ln)] += num('# Samples'); inst[(cur, ln)] += num('Instructions Executed'); srcs[(cur, ln)] = r[1]
    for k in KEYS: stall[(cur, ln)][k] += num(k)
tot = sum(agg.values()); toti = sum(inst.values())
print('total samples', tot, 'total warp instr', toti)
allst = collections.Counter()
for v in stall.values(): allst.update(v)
print('stalls', [(k, round(100 * v / max(tot, 1), 1)) for k, v in allst.most_common(8)])
for (f, ln), s in agg.most_common(top):
    st = [(k[6:], v) for k, v in stall[(f, ln)].most_common(3)]
    print(f"{f}:{ln:4d} {100*s/max(tot,1):5.1f}% inst {100*inst[(f,ln)]/max(toti,1):5.1f}%  {srcs[(f,ln)].strip()[:64]:64s} {st}")
