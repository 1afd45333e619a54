"""BASELINE config 5: one 32768-instance batch sharded by instance over 1/2/4/8 GPUs (strong scaling).
Run on the GPU box once per GPU count (gpurun --gpus N):
    N=1:  python bench.py --strong-total 32768 --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/strong_1.json
    N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N --strong-total 32768 --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/strong_N.json
then here:  python scripts/strong_scaling.py   ->  profiles/r2_strong_scaling.json"""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "gpurun_out", f"strong_{n}.json")
    if not os.path.exists(p):
        continue
    d = json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])
    rows.append({"n_gpus": n, "instances_per_gpu": d["config"]["instances_per_gpu"], "solves_per_s": d["value"], "ms_per_step": d["ms_per_step"],
                 "e2e_solves_per_s": d["e2e"]["value"], "clocks": d["clocks"]})
if rows:
    base = rows[0]["solves_per_s"] / rows[0]["n_gpus"]
    for r in rows:
        r["speedup_vs_1gpu"] = r["solves_per_s"] / (base * 1.0) if rows[0]["n_gpus"] == 1 else None
        r["efficiency"] = r["speedup_vs_1gpu"] / r["n_gpus"] if r["speedup_vs_1gpu"] else None
    out = {"workload": "BASELINE config 5: ONE batch of 32768 instances (config-2 generator), fixed 2x3 HSDDP iterations, sharded by instance over N GPUs of one box; device-timed, max over ranks",
           "rows": rows}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_strong_scaling.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
