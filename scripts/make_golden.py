"""Generate the golden vectors under tests/golden/ from the oracle (strict build: -O2 -ffp-contract=off) linked to the
reference's own generated C.  The reference ships no tests or golden vectors, so these are
produced here, in the build container where /root/reference is present, and committed.

    python scripts/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_specs, to_orc_problem
from oracle.oracle_py import OracleSolver, eval_fn

OUT = os.path.join(ROOT, "tests", "golden")


def solve_trace(spec, **opt):
    s = OracleSolver(to_orc_problem(spec, **opt), variant="strict")
    s.initialization()
    warm = {"x0_phase0": s.field(0, "X_NOM")[[0, 1, 40, -1]].tolist(), "u_phase0": s.field(0, "U_NOM")[[0, 1, 40]].tolist()}
    st = s.solve()
    modes, N, wb = s.schedule()
    fin = {"x_first": s.field(0, "X_NOM")[0].tolist(), "x_last_wb": s.field(3 if len(N) > 3 else 0, "X_NOM")[-1].tolist(),
           "u_phase0_k0_k10": s.field(0, "U_NOM")[[0, 10]].tolist(), "K_phase0_k0": s.field(0, "K")[0].tolist(),
           "du_phase0_k0": s.field(0, "DU")[0].tolist()}
    return {"status": st, "schedule": {"mode": modes, "N": N, "is_wb": wb}, "records": s.records(), "scalars": s.scalars(),
            "warmstart": warm, "final": fin, "phase_params": [s.phase_params(p) for p in range(s.n_phases)]}


def main():
    os.makedirs(OUT, exist_ok=True)
    g = {"about": "oracle (strict) traces; see scripts/make_golden.py", "cases": {}}
    g["cases"]["config1_default"] = solve_trace(make_specs(1, 1)[0])
    for i in (1, 2, 3):
        g["cases"][f"config2_inst{i}"] = solve_trace(make_specs(4, 2)[i])
    g["cases"]["config3_allwb_inst1"] = solve_trace(make_specs(2, 3)[1])
    for i in (0, 5):
        g["cases"][f"config4_inst{i}"] = solve_trace(make_specs(6, 4)[i])
    g["cases"]["config2_inst1_fixed_budget"] = solve_trace(make_specs(2, 2)[1], DDP_thresh=-1.0, AL_thresh=-1.0)
    json.dump(g, open(os.path.join(OUT, "solve_traces.json"), "w"), indent=1)
    # known-answer vectors of the generated functions at the default x0 (+ one perturbed state)
    kv = {}
    rng = np.random.default_rng(7)
    x = np.array(make_specs(2, 2)[1]["x0"]); u = np.array([3.0, -7.5, 11.0, -2.25])
    for name in ("Dyn_BS", "Dyn_FS", "Dyn_FL", "Dyn_BS_par", "Dyn_FS_par", "Dyn_FL_par"):
        kv[name] = {"in": [x.tolist(), u.tolist()], "out": [o.flatten(order="F").tolist() for o in eval_fn(name, x, u)]}
    for name in ("Imp_F", "Imp_B", "Imp_F_par", "Imp_B_par", "WB_FL1_terminal_constr", "WB_FL2_terminal_constr", "Jacob_F", "Jacob_B"):
        kv[name] = {"in": [x.tolist()], "out": [o.flatten(order="F").tolist() for o in eval_fn(name, x)]}
    xr = [0.1, -0.12, 0.05, 1.4, 0.2, -0.3]; fh = [0.3, -0.404, -0.1, -0.404]
    for ct in ([0.0, 1.0], [1.0, 0.0]):
        for name in ("FBDynamics", "FBDynamics_par"):
            kv[f"{name}_ct{int(ct[0])}{int(ct[1])}"] = {"fn": name, "in": [xr, u.tolist(), fh, ct],
                                                        "out": [o.flatten(order="F").tolist() for o in eval_fn(name, xr, u, fh, ct)]}
    json.dump(kv, open(os.path.join(OUT, "function_kat.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
