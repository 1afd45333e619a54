import sys, os, random, time
sys.path.insert(0,'/root/repo')
import mhpc_minimal_env_b200.codegen.gen_all as G
from mhpc_minimal_env_b200.codegen.sxdag import parse_casadi_c, CONST
from mhpc_minimal_env_b200.codegen.implicit import implicit_jacobian
from mhpc_minimal_env_b200.codegen.emit import reordered
from mhpc_minimal_env_b200.codegen.regalloc import allocate, pressure_order, spill_cost
fns = {}
for name in G.FUNCTIONS:
    path = os.path.join(G.DEFAULT_REF, "CasadiGen", "source", name + ".c")
    if os.path.exists(path): fns[name] = parse_casadi_c(path)
def devfn(name):
    primal = G.IMPLICIT_FUNCTIONS[name]
    return implicit_jacobian(fns[primal], fns[name], G.USE_LDLT, donor=fns[G.DONOR[primal]] if primal in G.DONOR else None)

def rand_order(fn, seed, w_kill=1.0, noise=0.0, la=0):
    rnd = random.Random(seed)
    nodes = fn.nodes; n = len(nodes)
    users = [[] for _ in range(n)]; npred = [0]*n
    for i, nd in enumerate(nodes):
        ps = set(p for p in (nd.a, nd.b) if p >= 0); npred[i] = len(ps)
        for p in ps: users[p].append(i)
    users = [sorted(set(u)) for u in users]; remaining = [len(u) for u in users]
    pri = [rnd.random() for _ in range(n)]
    def ops_of(i):
        nd = nodes[i]; return set(q for q in (nd.a, nd.b) if q >= 0 and nodes[q].op != CONST)
    def score(i):
        nd = nodes[i]
        if nd.op == CONST: return (-10, 0, 0)
        kills = sum(1 for p in ops_of(i) if remaining[p] == 1)
        base = (1 if users[i] else 0) - kills
        bonus = 0
        if la:
            for u in users[i]:
                if npred[u] == 1: bonus = max(bonus, sum(1 for p in ops_of(u) if remaining[p] == 1))
        return (base, -bonus, i + noise * n * pri[i])
    ready = set(i for i in range(n) if npred[i] == 0); order = []
    while ready:
        best = min(ready, key=score); ready.discard(best); order.append(best)
        for p in set(q for q in (nodes[best].a, nodes[best].b) if q >= 0): remaining[p] -= 1
        for u in users[best]:
            npred[u] -= 1
            if npred[u] == 0: ready.add(u)
    return order

for name in ("Dyn_BS_par", "Dyn_FL_par"):
    fn = devfn(name)
    base = {}
    for tag in ("index", "unlock"):
        st = allocate(reordered(fn, pressure_order(fn, tag)), 104, 24, 0, 108, 128, G.quad_map(name))[1]
        base[tag] = spill_cost(st)
    print(name, "baseline costs", base, flush=True)
    best = None
    t0 = time.time()
    for seed in range(12):
        for la in (0, 1):
            for noise in (0.02, 0.1, 0.5):
                o = rand_order(fn, seed, noise=noise, la=la)
                st = allocate(reordered(fn, o), 104, 24, 0, 108, 128, G.quad_map(name))[1]
                c = spill_cost(st)
                if best is None or c < best[0]:
                    best = (c, seed, la, noise, st["loads"], st["stores"], st["smem_loads"], st["smem_stores"])
                    print("   new best", best, round(time.time() - t0), "s", flush=True)

print("==== wide search", flush=True)
import itertools
res = {}
for name in ("Dyn_BS_par", "Dyn_FS_par", "Dyn_FL_par", "Imp_F_par", "Imp_B_par"):
    fn = devfn(name)
    best = None; t0 = time.time()
    for seed in range(40):
        for noise in (0.3, 1.0, 3.0):
            o = rand_order(fn, seed, noise=noise, la=1)
            st = allocate(reordered(fn, o), 108, 24, 0, 108, 128, G.quad_map(name))[1]
            c = spill_cost(st)
            if best is None or c < best[0]:
                best = (c, seed, noise, st["loads"], st["stores"], st["smem_loads"], st["smem_stores"])
    res[name] = best
    print(name, best, round(time.time() - t0), "s", flush=True)
print(res)
