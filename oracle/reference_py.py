"""TEST INFRASTRUCTURE (oracle) - ctypes wrapper around oracle/_ref/libmhpc_reference_*.so: the reference's own,
unmodified solver sources compiled against oracle/eigen_shim (see oracle/Makefile, target `reference`).
Used by tests/ to pin the restatement (hsddp_oracle.cpp) and by bench.py as the CPU reference arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_OPT = ("alpha", "gamma", "update_penalty", "update_relax", "update_regularization", "update_ReB", "max_DDP_iter",
        "max_AL_iter", "DDP_thresh", "AL_thresh")
_DEF = dict(alpha=0.1, gamma=0.01, update_penalty=8, update_relax=0.1, update_regularization=2, update_ReB=7,
            max_DDP_iter=3, max_AL_iter=2, DDP_thresh=1e-3, AL_thresh=1e-3, AL_active=1, ReB_active=1, smooth_active=0)
_libs = {}


def available(variant: str = "strict") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libmhpc_reference_{variant}.so")) or \
        os.path.isdir("/root/reference/HSDDPSolver/source")


def lib(variant: str = "strict"):
    if variant in _libs:
        return _libs[variant]
    so = os.path.join(HERE, "_ref", f"libmhpc_reference_{variant}.so")
    if not os.path.exists(so):
        if not os.path.isdir("/root/reference/HSDDPSolver/source"):
            raise RuntimeError("oracle/_ref/libmhpc_reference is not built and /root/reference is absent")
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "reference"])
    L = C.CDLL(so)
    L.ref_create.restype = C.c_void_p
    _libs[variant] = L
    return L


class quiet_stdout:
    """the reference prints its DEBUG trace with printf (MHPC_CPPTypes.h:20); silence fd 1 while it runs"""
    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class ReferenceSolver:
    def __init__(self, spec: dict, variant: str = "strict", **opt):
        self.L = lib(variant)
        o = dict(_DEF)
        o.update(opt)
        opt10 = (C.c_double * 10)(*[o[k] for k in _OPT])
        fl = (C.c_int * 3)(int(o["AL_active"]), int(o["ReB_active"]), int(o["smooth_active"]))
        gm = (C.c_int * 4)(*spec["gait_mode"])
        gt = (C.c_float * 4)(*spec["gait_timing"])
        x0 = (C.c_double * 14)(*spec["x0"])
        self.h = C.c_void_p(self.L.ref_create(opt10, fl, spec["n_wbphase"], spec["n_fbphase"], C.c_float(spec["dt_wb"]),
                                              C.c_float(spec["dt_fb"]), spec["cmode"], C.c_float(spec["vel"]),
                                              C.c_float(spec["height"]), gm, gt, x0))

    def initialization(self): self.L.ref_initialization(self.h)
    def update_problem(self): self.L.ref_update_problem(self.h)

    def set_x0(self, x0):
        a = (C.c_double * 14)(*[float(v) for v in x0])
        self.L.ref_set_x0(self.h, a)

    def solve(self):
        with quiet_stdout():
            self.L.ref_solve(self.h)

    def forward_sweep(self, eps): self.L.ref_forward_sweep(self.h, C.c_double(eps))
    def update_nominal_trajectory(self): self.L.ref_update_nominal(self.h)
    def backward_sweep(self, reg) -> bool: return bool(self.L.ref_backward_sweep(self.h, C.c_double(reg)))

    def forward_iteration(self) -> int:
        with quiet_stdout():
            return self.L.ref_forward_iteration(self.h)

    def print_debugInfo(self, directory: str):
        """the reference's own print_debugInfo: writes state.txt / control.txt / gradient.txt / cost.txt into `directory`"""
        with quiet_stdout():
            rc = self.L.ref_print_debugInfo(self.h, str(directory).encode())
        assert rc == 0

    def scalars(self):
        o = (C.c_double * 3)()
        self.L.ref_scalars(self.h, o)
        return {"cost": o[0], "exp_cost_change": o[1], "viol": o[2]}

    def schedule(self):
        n = self.L.ref_n_phases(self.h)
        m, N = (C.c_int * n)(), (C.c_int * n)()
        self.L.ref_schedule(self.h, m, N)
        return list(m), list(N)

    def field(self, phase: int, name: str) -> np.ndarray:
        which = {"X_NOM": 0, "U_NOM": 1, "Y_NOM": 2, "K": 3, "DU": 4, "G": 5, "H": 6, "QX": 7, "QU": 8, "QXX": 9, "QUX": 10, "QUU": 11}[name]
        dim = self.L.ref_get(self.h, phase, which, None)
        _, N = self.schedule()
        out = np.zeros((N[phase], dim))
        self.L.ref_get(self.h, phase, which, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out
