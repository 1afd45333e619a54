"""Build-time driver: reference CasADi SX C  ->  device bodies + metadata.

    python -m mhpc_minimal_env_b200.codegen.gen_all [--ref /root/reference] [--out DIR]

Reads ``<ref>/CasadiGen/source/<F>.c`` for the 16 functions on (or feeding) the
hot path (a2-a7, a11 and the warm-start Jacobians) and
writes into ``mhpc_minimal_env_b200/csrc/gen`` (git-ignored build artefacts):

    <F>.inc         macro-parameterised straight-line body (see emit.py)
    gen_check.c     the same DAGs as plain C, for the bit-exactness test on the CPU
    gen_meta.h      sparsity / structure tables the kernels static_assert against
    gen_meta.json   the same for the Python tests

The structure checks at the bottom pin every assumption the hand-written kernels
make about the generated sparsity (identity blocks, empty columns, constant
entries); generation fails loudly if the reference expressions ever change shape.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

from .emit import critical_path_order, emit_c, emit_inc, reordered
from .implicit import implicit_jacobian, structured_primal
from .regalloc import pressure_order, spill_cost, allocate as ra_allocate, emit_c_regalloc, emit_inc_regalloc
from .sxdag import CONST, SXFunction, parse_casadi_c, share_reciprocals

FUNCTIONS = [
    "Dyn_BS", "Dyn_FS", "Dyn_FL",
    "Dyn_BS_par", "Dyn_FS_par", "Dyn_FL_par",
    "Imp_F", "Imp_B", "Imp_F_par", "Imp_B_par",
    "FBDynamics", "FBDynamics_par",
    "WB_FL1_terminal_constr", "WB_FL2_terminal_constr",
    "Jacob_F", "Jacob_B",
    # link-point Jacobians: included by the reference (CasadiGen.h:22-26) but never called (PlanarQuadruped.h:72 is an empty
    # stub); generated for completeness and reachable through mhpc_eval_function only
    "Link0Jacobian", "Link1Jacobian", "Link2Jacobian", "Link3Jacobian", "Link4Jacobian",
]

# the large Jacobian bodies additionally get a generator-side register allocation (codegen/regalloc.py)
REGALLOC_FUNCTIONS = ["Dyn_BS_par", "Dyn_FS_par", "Dyn_FL_par", "Imp_F_par", "Imp_B_par"]
# dynamics Jacobians re-derived by implicit differentiation of the primal's linear solve (codegen/implicit.py); the
# flat CasADi DAGs stay the CPU check (gen_/genra_) and MHPC_IMPLICIT_JAC=0 puts them back on the device
IMPLICIT_FUNCTIONS = {"Dyn_BS_par": "Dyn_BS", "Dyn_FS_par": "Dyn_FS", "Dyn_FL_par": "Dyn_FL", "Imp_F_par": "Imp_F", "Imp_B_par": "Imp_B"}
# the impact maps' right-hand side has two structurally zero rows; their system matrix is taken from the stance dynamics of
# the same foot (codegen/implicit.py: identify_system(donor=...))
DONOR = {"Imp_F": "Dyn_FS", "Imp_B": "Dyn_BS"}
IMPLICIT_JAC = os.environ.get("MHPC_IMPLICIT_JAC", "1") == "1"
# the primal dynamics with their symbolic Gram-Schmidt QR replaced by an LDL^T solve of the same system (rollouts)
STRUCT_PRIMAL = os.environ.get("MHPC_STRUCT_PRIMAL", "1") == "1"
USE_LDLT = os.environ.get("MHPC_LDLT", "1") == "1"
CP_ORDER = os.environ.get("MHPC_CP_ORDER", "0") == "1"      # primal dynamics emitted along their critical path: neutral (ptxas -O3 re-schedules)
JAC_CP = os.environ.get("MHPC_JAC_CP", "0") == "1"          # same for the Jacobian bodies, which ptxas -O1 keeps in the generator's order
RA_CAP = int(os.environ.get("MHPC_RA_CAP", "108"))
RA_LOOKAHEAD = int(os.environ.get("MHPC_RA_LOOKAHEAD", "24"))
RA_WINDOW = int(os.environ.get("MHPC_RA_WINDOW", "0"))
SHARE_RCP = os.environ.get("MHPC_SHARE_RCP", "1") == "1"
RA_FAR = int(os.environ.get("MHPC_RA_FAR", "128"))           # prefetch distance for operands in local memory (0: same as lookahead)
RA_SMEM = int(os.environ.get("MHPC_RA_SMEM", "108"))      # per-thread shared-memory slots of the second spill level
RA_ORDER = os.environ.get("MHPC_RA_ORDER", "auto")        # evaluation order of the Jacobian bodies: given | index | unlock | tuned | auto (cheapest of them)
# (seed, noise) of the best register-pressure list schedule found by an offline search over 700 seeds x 2-3 noise levels at the default
# allocator settings (scripts/order_search.py; cost = spill wavefronts, regalloc.spill_cost): 3122 -> 868 for Dyn_BS_par, the far level
# of Dyn_FL_par and of the impact maps empty.  `auto` still compares them with the untuned orders under the settings in force.
TUNED_ORDER = {"Dyn_BS_par": (394, 100.0), "Dyn_FS_par": (558, 100.0), "Dyn_FL_par": (523, 100.0), "Imp_F_par": (21, 0.3), "Imp_B_par": (3, 0.3)}
RA_STAGE = int(os.environ.get("MHPC_RA_STAGE", "0"))      # per-thread staging slots of the asynchronous far level (0: synchronous local-memory array)

QUAD_OUT = os.environ.get("MHPC_QUAD_OUT", "1") == "1"


def wb_record_element(st):
    """Knot-record element (csrc/common.cuh: R_AB, R_BB, R_CB, R_DB) a Dyn_*_par output non-zero lands in, or None for
    the rows the kernels do not store (the identity block of A).  k_lin.cu static_asserts the same mapping."""
    if st.out == 0:
        return (st.col - 2) * 7 + (st.row - 7) if st.row >= 7 else None
    if st.out == 1:
        return 84 + st.col * 7 + (st.row - 7)
    if st.out == 2:
        return 112 + (st.col - 2) * 2 + (st.row & 1)
    return 136 + st.col * 2 + (st.row & 1)


def imp_record_element(st):
    """Element of a WB terminal-knot record (R_AB block = Px[7:14, 2:14]) an Imp_*_par non-zero lands in."""
    return (st.col - 2) * 7 + (st.row - 7) if st.row >= 7 else None


def quad_map(name):
    if not (QUAD_OUT and IMPLICIT_JAC and name in IMPLICIT_FUNCTIONS):
        return None
    return imp_record_element if name.startswith("Imp_") else wb_record_element


HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_OUT = os.path.normpath(os.path.join(HERE, "..", "csrc", "gen"))
DEFAULT_REF = "/root/reference"


def _const_stores(fn: SXFunction, out: int):
    return {(s.row, s.col): fn.nodes[s.node].val for s in fn.stores
            if s.out == out and fn.nodes[s.node].op == CONST}


def _coords(fn: SXFunction, out: int):
    return sorted((s.row, s.col) for s in fn.stores if s.out == out)


def check_structure(fns: dict) -> None:
    """Assumptions baked into csrc/*.cuh (cited there)."""
    wb_ac = sorted([(i, 7 + i) for i in range(7)] + [(r, c) for r in range(7, 14) for c in range(2, 14)])
    wb_bc = sorted((r, c) for r in range(7, 14) for c in range(4))
    for name, yrows in (("Dyn_BS_par", (2, 3)), ("Dyn_FS_par", (0, 1)), ("Dyn_FL_par", ())):
        f = fns[name]
        assert _coords(f, 0) == wb_ac, name
        assert _const_stores(f, 0) == {(i, 7 + i): 1.0 for i in range(7)}, name
        assert _coords(f, 1) == wb_bc, name
        assert _coords(f, 2) == sorted((r, c) for r in yrows for c in range(2, 14)), name
        assert _coords(f, 3) == sorted((r, c) for r in yrows for c in range(4)), name
    px = sorted([(i, i) for i in range(7)] + [(r, c) for r in range(7, 14) for c in range(2, 14)])
    for name in ("Imp_F_par", "Imp_B_par"):
        f = fns[name]
        assert _coords(f, 0) == px, name
        assert _const_stores(f, 0) == {(i, i): 1.0 for i in range(7)}, name
    f = fns["FBDynamics_par"]
    assert _coords(f, 0) == sorted([(0, 3), (1, 4), (2, 5), (5, 0), (5, 1)])
    assert _const_stores(f, 0) == {(0, 3): 1.0, (1, 4): 1.0, (2, 5): 1.0}
    assert _coords(f, 1) == sorted([(3, 0), (5, 0), (4, 1), (5, 1), (3, 2), (5, 2), (4, 3), (5, 3)])
    for name in ("WB_FL1_terminal_constr", "WB_FL2_terminal_constr"):
        f = fns[name]
        assert _coords(f, 0) == [(0, 0)]
        hx = _coords(f, 1)
        hxx = _coords(f, 2)
        assert len(hx) == 4 and all(r == 0 and 1 <= c <= 6 for r, c in hx), (name, hx)
        assert len(hxx) == 9 and all(2 <= r <= 6 and 2 <= c <= 6 for r, c in hxx), (name, hxx)
    for name in ("Dyn_BS", "Dyn_FS", "Dyn_FL", "Imp_F", "Imp_B"):
        f = fns[name]
        assert _coords(f, 0) == [(i, 0) for i in range(14)] and _coords(f, 1) == [(i, 0) for i in range(4)]


def meta_of(fn: SXFunction) -> dict:
    outs = []
    for k, sp in enumerate(fn.out_sp):
        st = sorted((s for s in fn.stores if s.out == k), key=lambda s: s.nz)
        outs.append({
            "nrow": sp.nrow, "ncol": sp.ncol, "nnz": sp.nnz,
            "row": [s.row for s in st], "col": [s.col for s in st],
            "const": [fn.nodes[s.node].val if fn.nodes[s.node].op == CONST else None for s in st],
        })
    return {"name": fn.name, "in_sizes": fn.in_sizes, "nodes": len(fn.nodes),
            "flops": fn.flops(), "census": fn.op_census(), "outputs": outs}


def emit_meta_h(metas: dict) -> str:
    L = ["// GENERATED by mhpc_minimal_env_b200/codegen/gen_all.py. Do not edit.",
         "#pragma once", "namespace mhpc_gen {"]
    for name, m in metas.items():
        L.append(f"// {name}: nodes {m['nodes']}, flops {m['flops']}")
        L.append(f"constexpr int {name}_FLOPS = {m['flops']};")
        for k, o in enumerate(m["outputs"]):
            n = max(o["nnz"], 1)
            rows = ", ".join(str(r) for r in o["row"]) or "0"
            cols = ", ".join(str(c) for c in o["col"]) or "0"
            L.append(f"constexpr int {name}_O{k}_NNZ = {o['nnz']};")
            L.append(f"constexpr int {name}_O{k}_ROW[{n}] = {{{rows}}};")
            L.append(f"constexpr int {name}_O{k}_COL[{n}] = {{{cols}}};")
    L.append("}  // namespace mhpc_gen")
    return "\n".join(L) + "\n"


def generate(ref: str = DEFAULT_REF, out: str = DEFAULT_OUT, verbose: bool = True) -> dict:
    src_dir = os.path.join(ref, "CasadiGen", "source")
    os.makedirs(out, exist_ok=True)
    fns, metas = {}, {}
    h = hashlib.sha256()
    for name in FUNCTIONS:
        path = os.path.join(src_dir, name + ".c")
        with open(path, "rb") as fh:
            h.update(fh.read())
        fn = parse_casadi_c(path)
        assert fn.name == name
        fns[name] = fn
        metas[name] = meta_of(fn)
    check_structure(fns)

    def write(rel, text):
        p = os.path.join(out, rel)
        old = None
        if os.path.exists(p):
            with open(p) as fh:
                old = fh.read()
        if old != text:                      # keep mtimes stable for make/ninja
            with open(p, "w") as fh:
                fh.write(text)

    # device bodies: shared reciprocals (<= 1 ulp per quotient); the exact-division DAGs stay the CPU check below
    dev = {name: (share_reciprocals(fn) if SHARE_RCP else fn) for name, fn in fns.items()}
    restated = []
    if IMPLICIT_JAC:
        for name, primal in IMPLICIT_FUNCTIONS.items():
            dev[name] = implicit_jacobian(fns[primal], fns[name], USE_LDLT, donor=fns[DONOR[primal]] if primal in DONOR else None)
            restated.append(name)
    if STRUCT_PRIMAL:
        for primal in IMPLICIT_FUNCTIONS.values():
            dev[primal] = structured_primal(fns[primal], USE_LDLT, donor=fns[DONOR[primal]] if primal in DONOR else None)
            restated.append(primal)
    for name, fn in dev.items():
        # the rollout's dynamics are emitted along their critical path (same DAG, same operations; emit.critical_path_order)
        write(name + ".inc", emit_inc(fn, critical_path_order(fn) if (CP_ORDER and name in IMPLICIT_FUNCTIONS.values()) else None))
    if JAC_CP:
        for name in IMPLICIT_FUNCTIONS:
            dev[name] = reordered(dev[name], critical_path_order(dev[name]))
    ra_stats = {}
    if RA_ORDER != "given":
        # evaluation order of the Jacobian bodies: of the given order and two register-pressure list schedules, the one whose
        # allocation moves the fewest spill wavefronts (same DAG, same operations, bit-identical results)
        for name in REGALLOC_FUNCTIONS:
            best, best_cost, best_tag = dev[name], None, "given"
            for tag in (("given", "index", "unlock", "tuned") if RA_ORDER == "auto" else (RA_ORDER,)):
                if tag == "tuned" and name not in TUNED_ORDER:
                    continue
                cand = (dev[name] if tag == "given" else
                        reordered(dev[name], pressure_order(dev[name], "unlock", noise=TUNED_ORDER[name][1], seed=TUNED_ORDER[name][0])) if tag == "tuned" else
                        reordered(dev[name], pressure_order(dev[name], tag)))
                cost = spill_cost(ra_allocate(cand, RA_CAP, RA_LOOKAHEAD, RA_WINDOW, RA_SMEM, RA_FAR, quad_map(name), RA_STAGE)[1])
                if best_cost is None or cost < best_cost:
                    best, best_cost, best_tag = cand, cost, tag
            dev[name] = best
            ra_stats.setdefault("_order", {})[name] = best_tag
    order_tags = ra_stats.pop("_order", {})
    for name in REGALLOC_FUNCTIONS:
        text, st = emit_inc_regalloc(dev[name], RA_CAP, RA_LOOKAHEAD, RA_WINDOW, RA_SMEM, RA_FAR, quad_map(name), RA_STAGE)
        st["order"] = order_tags.get(name, "given")
        write(name + "_ra.inc", text)
        ra_stats[name] = st
    write("gen_check.c", "".join(emit_c(fn, "gen_" + name) for name, fn in fns.items()) +
          "".join(emit_c_regalloc(fns[name], "genra_" + name, RA_CAP, RA_LOOKAHEAD, RA_WINDOW, RA_SMEM, RA_FAR, None, RA_STAGE) for name in REGALLOC_FUNCTIONS) +
          "".join(emit_c_regalloc(dev[name], "gendev_" + name, RA_CAP, RA_LOOKAHEAD, RA_WINDOW, RA_SMEM, RA_FAR, quad_map(name), RA_STAGE) for name in REGALLOC_FUNCTIONS) +
          "".join(emit_c(dev[name], "gendev_" + name) for name in restated if name not in REGALLOC_FUNCTIONS))
    write("gen_meta.h", emit_meta_h(metas) + f"#define MHPC_RA_SMEM_SLOTS {RA_SMEM}\n#define MHPC_RA_STAGE_SLOTS {RA_STAGE}\n"
          + f"#define MHPC_RA_FAR_SLOTS {max([st.get('slots', 0) for st in ra_stats.values()] + [1])}\n")
    metas["_source_sha256"] = h.hexdigest()
    metas["_regalloc"] = {"cap": RA_CAP, "lookahead": RA_LOOKAHEAD, "window": RA_WINDOW, "share_reciprocals": SHARE_RCP, "smem_slots": RA_SMEM, "stage_slots": RA_STAGE, "far_lookahead": RA_FAR, "stats": ra_stats,
                         "implicit_jacobians": sorted(restated), "ldlt": USE_LDLT,
                         "device_nodes": {name: len(dev[name].nodes) for name in FUNCTIONS},
                         "device_flops": {name: dev[name].flops() for name in FUNCTIONS}}
    write("gen_meta.json", json.dumps(metas, indent=1))
    if verbose:
        for name in FUNCTIONS:
            m = metas[name]
            print(f"[codegen] {name:<24} nodes {m['nodes']:>6}  flops {m['flops']:>6}")
    return metas


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF)
    ap.add_argument("--out", default=DEFAULT_OUT)
    a = ap.parse_args(argv)
    if not os.path.isdir(os.path.join(a.ref, "CasadiGen", "source")):
        print(f"[codegen] reference tree not found at {a.ref}", file=sys.stderr)
        return 2
    generate(a.ref, a.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
