"""Build driver: code generation + nvcc (sm_100a) -> mhpc_minimal_env_b200/libmhpc_b200.so.

    python -m mhpc_minimal_env_b200.build [--ref /root/reference] [--force]

1. If the reference tree is present, (re)generate the device bodies csrc/gen/*.inc from the CasADi-emitted C
   (mhpc_minimal_env_b200/codegen).  They are build artefacts: git-ignored, never committed.
2. Compile every translation unit for sm_100a only (-gencode arch=compute_100a,code=sm_100a -lineinfo), in
   parallel, and link the C-ABI shared library declared in include/mhpc_b200.h.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmhpc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]

# resident 128-thread blocks per SM the Jacobian kernels are compiled for (register cap = 65536 / (128 * LIN_MINBLOCKS))
LIN_MINBLOCKS = int(os.environ.get("MHPC_LIN_MINBLOCKS", "2"))

LIN_EXTRA = os.environ.get("MHPC_LIN_EXTRA", "").split()   # dev: extra nvcc flags for the Jacobian kernel classes
BWS_EXTRA = os.environ.get("MHPC_BWS_EXTRA", "").split()   # dev: extra nvcc flags for the backward sweep
# ptxas level for the Jacobian kernels: -O1 keeps the generator's instruction order and register allocation (the bodies
# come pre-allocated by codegen/regalloc.py); higher levels re-schedule them
LIN_OPT = os.environ.get("MHPC_LIN_OPT", "O1")


def _lin_flags(c):
    if c >= 5:
        return []
    return ["-Xptxas", "-" + LIN_OPT, f"-DLIN_MINBLOCKS={LIN_MINBLOCKS}"] + LIN_EXTRA


# (source, object name, extra defines)
UNITS = [("capi.cu", "capi.o", []), ("k_misc.cu", "k_misc.o", []), ("k_rollout.cu", "k_rollout.o", []),
         ("k_bws.cu", "k_bws.o", BWS_EXTRA), ("k_bws16.cu", "k_bws16.o", BWS_EXTRA), ("k_bwsT.cu", "k_bwsT.o", BWS_EXTRA), ("k_eval.cu", "k_eval.o", [])] + \
        [("k_lin.cu", f"k_lin_{c}.o", [f"-DLIN_CLS={c}"] + _lin_flags(c)) for c in range(6)]
# ptxas -O3 degenerates on the 26k-node Jacobian bodies (32 registers, ~100 KB of spill stores per thread);
# at -O1 it keeps the generator's instruction order and allocates 255 registers (~24 KB of spill stores).

GEN_FILES = ["Dyn_BS", "Dyn_FS", "Dyn_FL", "Dyn_BS_par", "Dyn_FS_par", "Dyn_FL_par", "Imp_F", "Imp_B", "Imp_F_par",
             "Imp_B_par", "FBDynamics", "FBDynamics_par", "WB_FL1_terminal_constr", "WB_FL2_terminal_constr",
             "Jacob_F", "Jacob_B", "Link0Jacobian", "Link1Jacobian", "Link2Jacobian", "Link3Jacobian", "Link4Jacobian"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(p.encode())
            h.update(fh.read())
    return h.hexdigest()


def generate(ref: str, verbose: bool = True) -> None:
    have_gen = all(os.path.exists(os.path.join(GEN, f + ".inc")) for f in GEN_FILES)
    if os.path.isdir(os.path.join(ref, "CasadiGen", "source")):
        sys.path.insert(0, os.path.dirname(HERE))
        from mhpc_minimal_env_b200.codegen.gen_all import generate as gen
        gen(ref, GEN, verbose=verbose)
    elif not have_gen:
        raise RuntimeError(f"csrc/gen is empty and the reference tree is not at {ref}: cannot generate device code")


def _compile(src, obj, defs, log_dir):
    out = os.path.join(OBJ, obj)
    cmd = [NVCC, *ARCH, *FLAGS, *defs, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", out]
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(log_dir, obj + ".log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    return obj, r.returncode, time.time() - t, r.stderr


def build(ref: str = "/root/reference", force: bool = False, verbose: bool = True) -> str:
    generate(ref, verbose=False)
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(GEN, f) for f in os.listdir(GEN) if f.endswith(".inc")]
    srcs += [os.path.join(HERE, "..", "include", "mhpc_b200.h"), os.path.abspath(__file__)]
    stamp = os.path.join(OBJ, "stamp.txt")
    dig = _digest(srcs) + f":mb{LIN_MINBLOCKS}:{LIN_OPT}:{os.environ.get('MHPC_RA_CAP', '')}:{os.environ.get('MHPC_RA_LOOKAHEAD', '')}:{os.environ.get('MHPC_RA_WINDOW', '')}:{os.environ.get('MHPC_SHARE_RCP', '')}:{os.environ.get('MHPC_RA_SMEM', '')}:{os.environ.get('MHPC_RA_FAR', '')}:{os.environ.get('MHPC_RA_STAGE', '')}:{os.environ.get('MHPC_RA_ORDER', '')}:{os.environ.get('MHPC_RA_IN_SMEM', '')}:{os.environ.get('MHPC_LIN_EXTRA', '')}:{os.environ.get('MHPC_BWS_EXTRA', '')}"
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        if verbose:
            print(f"[build] up to date: {LIB}")
        return LIB
    t0 = time.time()
    with cf.ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        futs = [ex.submit(_compile, s, o, d, OBJ) for s, o, d in UNITS]
        ok = True
        for f in cf.as_completed(futs):
            obj, rc, dt, err = f.result()
            if verbose:
                print(f"[build] {obj:<16} rc={rc} {dt:6.1f}s")
            if rc != 0:
                ok = False
                sys.stderr.write(err[-4000:])
    if not ok:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OBJ, o) for _, o, _ in UNITS]
    subprocess.check_call([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-cudart", "shared"])
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print(f"[build] linked {LIB} in {time.time() - t0:.1f}s")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(a.ref, a.force)
