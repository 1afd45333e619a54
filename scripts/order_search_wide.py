import sys, os, time
sys.path.insert(0,'/root/repo')
exec(open('/root/repo/scripts/order_search.py').read().split('for name in ("Dyn_BS_par", "Dyn_FL_par"):')[0])
name = sys.argv[1]
fn = devfn(name)
best = None; t0 = time.time()
for seed in range(int(sys.argv[2]), int(sys.argv[3])):
    for noise in (10.0, 100.0):
        o = pressure_order(fn, "unlock", noise=noise, seed=seed)
        st = allocate(reordered(fn, o), 108, 24, 0, 108, 128, G.quad_map(name))[1]
        c = spill_cost(st)
        if best is None or c < best[0]:
            best = (c, seed, noise, st["loads"], st["stores"], st["smem_loads"], st["smem_stores"])
            print(name, "new best", best, round(time.time() - t0), "s", flush=True)
print(name, "FINAL", best, flush=True)
