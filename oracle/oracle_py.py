"""TEST INFRASTRUCTURE (oracle) - ctypes wrapper around oracle/build/libmhpc_oracle_*.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product (mhpc_minimal_env_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FIELDS = ["X_ACT", "U_ACT", "Y_ACT", "X_NOM", "U_NOM", "Y_NOM", "X_REF", "U_REF", "Y_REF",
          "A", "B", "C", "D", "LX", "LU", "LY", "LXX", "LUU", "LYY", "K", "DU", "G", "H", "PHIX", "PHIXX",
          "QX", "QU", "QXX", "QUX", "QUU"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}

DEFAULT_X0 = [0.0927, -0.1093, -0.1542, 1.0957, -2.2033, 0.9742, -1.7098,
              0.9011, 0.2756, 0.7333, 0.0446, 0.0009, 1.3219, 2.7346]


class OrcProblem(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("alpha", "gamma", "update_penalty", "update_relax",
                                          "update_regularization", "update_ReB", "max_DDP_iter",
                                          "max_AL_iter", "DDP_thresh", "AL_thresh")] + [
        ("AL_active", C.c_int), ("ReB_active", C.c_int), ("smooth_active", C.c_int),
        ("n_wbphase", C.c_int), ("n_fbphase", C.c_int), ("dt_wb", C.c_float), ("dt_fb", C.c_float),
        ("cmode", C.c_int), ("vel", C.c_float), ("height", C.c_float),
        ("gait_mode", C.c_int * 4), ("gait_timing", C.c_float * 4), ("x0", C.c_double * 14)]


def default_problem(**kw) -> OrcProblem:
    """test_main.cpp:16-32 + struct defaults (MHPC_CompoundTypes.h:196-251, Gait.h:21-26)."""
    p = OrcProblem()
    p.alpha, p.gamma, p.update_penalty, p.update_relax = 0.1, 0.01, 8, 0.1
    p.update_regularization, p.update_ReB = 2, 7
    p.max_DDP_iter, p.max_AL_iter, p.DDP_thresh, p.AL_thresh = 3, 2, 1e-3, 1e-3
    p.AL_active, p.ReB_active, p.smooth_active = 1, 1, 0
    p.n_wbphase, p.n_fbphase, p.dt_wb, p.dt_fb, p.cmode = 4, 4, 0.001, 0.001, 1
    p.vel, p.height = 1.5, 0.0
    p.gait_mode[:] = [1, 2, 3, 4]
    p.gait_timing[:] = [0.08, 0.1, 0.08, 0.1]
    p.x0[:] = DEFAULT_X0
    for k, v in kw.items():
        if k in ("gait_mode", "gait_timing", "x0"):
            getattr(p, k)[:] = list(v)
        else:
            setattr(p, k, v)
    return p


def build(variant: str = "strict") -> str:
    """Build (if needed) and return the path of the oracle library. Needs /root/reference only if
    oracle/_ref is not already built."""
    so = os.path.join(HERE, "build", f"libmhpc_oracle_{variant}.so")
    ref = os.path.join(HERE, "_ref", f"libcasadigen_ref_{variant}.so")
    if not (os.path.exists(so) and os.path.exists(ref)) or \
            os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "hsddp_oracle.cpp")):
        if not os.path.exists(ref) and not os.path.isdir("/root/reference/CasadiGen/source"):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "oracle"])
    return so


_libs = {}


def lib(variant: str = "strict"):
    if variant in _libs:
        return _libs[variant]
    L = C.CDLL(build(variant))
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(OrcProblem)]
    L.orc_set_x0.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.orc_set_x0.restype = None
    for fn in ("orc_destroy", "orc_initialization", "orc_forward_sweep_partials_only",
               "orc_update_nominal_trajectory", "orc_update_AL_ReB_param", "orc_update_problem"):
        getattr(L, fn).argtypes = [C.c_void_p]
        getattr(L, fn).restype = None
    for fn in ("orc_forward_sweep", "orc_forward_sweep_dynamics_only", "orc_set_update_penalty"):
        getattr(L, fn).argtypes = [C.c_void_p, C.c_double]
        getattr(L, fn).restype = None
    L.orc_backward_sweep.argtypes = [C.c_void_p, C.c_double]
    L.orc_set_ReB_active.argtypes = [C.c_void_p, C.c_int]
    for fn in ("orc_forward_iteration", "orc_solve", "orc_n_phases", "orc_n_records"):
        getattr(L, fn).argtypes = [C.c_void_p]
    L.orc_get_schedule.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    L.orc_get_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.orc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.orc_get_phase_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.orc_get_Px.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.orc_get_record.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.orc_eval.argtypes = [C.c_char_p] + [C.POINTER(C.c_double)] * 8
    _libs[variant] = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleSolver:
    """Mirror of the reference's MHPCLocomotion<double> / MultiPhaseDDP<double> method set."""

    def __init__(self, problem: OrcProblem | None = None, variant: str = "strict"):
        self.L = lib(variant)
        self.problem = problem if problem is not None else default_problem()
        self.h = C.c_void_p(self.L.orc_create(C.byref(self.problem)))

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def initialization(self): self.L.orc_initialization(self.h)
    def update_problem(self): self.L.orc_update_problem(self.h)

    def set_x0(self, x0):
        a = (C.c_double * 14)(*[float(v) for v in x0])
        self.L.orc_set_x0(self.h, a)

    def forward_sweep(self, eps): self.L.orc_forward_sweep(self.h, eps)
    def forward_sweep_dynamics_only(self, eps): self.L.orc_forward_sweep_dynamics_only(self.h, eps)
    def forward_sweep_partials_only(self): self.L.orc_forward_sweep_partials_only(self.h)
    def backward_sweep(self, reg) -> bool: return bool(self.L.orc_backward_sweep(self.h, reg))
    def forward_iteration(self) -> int: return self.L.orc_forward_iteration(self.h)
    def update_nominal_trajectory(self): self.L.orc_update_nominal_trajectory(self.h)
    def update_AL_ReB_param(self): self.L.orc_update_AL_ReB_param(self.h)
    def set_ReB_active(self, on): self.L.orc_set_ReB_active(self.h, int(on))
    def set_update_penalty(self, v): self.L.orc_set_update_penalty(self.h, float(v))
    def solve(self) -> int: return self.L.orc_solve(self.h)

    @property
    def n_phases(self): return self.L.orc_n_phases(self.h)

    def schedule(self):
        n = self.n_phases
        m, N, w = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        self.L.orc_get_schedule(self.h, m, N, w)
        return list(m), list(N), [bool(x) for x in w]

    def scalars(self):
        o = (C.c_double * 4)()
        self.L.orc_get_scalars(self.h, o)
        return {"cost": o[0], "exp_cost_change": o[1], "viol": o[2], "last_eps": o[3]}

    def field(self, phase: int, name: str) -> np.ndarray:
        """[N, dim] for per-knot fields (matrices flattened column-major), [dim] for PHIX/PHIXX."""
        fid = FIELD_ID[name]
        dim = self.L.orc_get_field(self.h, phase, fid, None)
        _, N, _ = self.schedule()
        if name in ("PHIX", "PHIXX"):
            out = np.zeros(dim)
        else:
            out = np.zeros((N[phase], dim))
        self.L.orc_get_field(self.h, phase, fid, _dp(out))
        return out

    def phase_params(self, phase: int):
        o = (C.c_double * 9)()
        self.L.orc_get_phase_params(self.h, phase, o)
        return {"sigma": o[0], "lambda": o[1], "delta": [o[2], o[3], o[4]], "eps": [o[5], o[6], o[7]], "h": o[8]}

    def Px(self, phase: int) -> np.ndarray:
        o = np.zeros(196)
        self.L.orc_get_Px(self.h, phase, _dp(o))
        return o.reshape(14, 14, order="F")

    def records(self):
        out = []
        for i in range(self.L.orc_n_records(self.h)):
            a, d = (C.c_int * 4)(), (C.c_double * 6)()
            self.L.orc_get_record(self.h, i, a, d)
            out.append({"iter_AL": a[0], "iter_DDP": a[1], "n_bws": a[2], "n_fit": a[3], "reg": d[0],
                        "eps": d[1], "cost": d[2], "exp_cost_change": d[3], "viol": d[4], "cost_prev": d[5]})
        return out


_SHAPES = {  # name -> (n_in, in sizes, out shapes)
    "Dyn_BS": ([14, 4], [(14, 1), (4, 1)]), "Dyn_FS": ([14, 4], [(14, 1), (4, 1)]), "Dyn_FL": ([14, 4], [(14, 1), (4, 1)]),
    "Dyn_BS_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]),
    "Dyn_FS_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]),
    "Dyn_FL_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]),
    "Imp_F": ([14], [(14, 1), (4, 1)]), "Imp_B": ([14], [(14, 1), (4, 1)]),
    "Imp_F_par": ([14], [(14, 14)]), "Imp_B_par": ([14], [(14, 14)]),
    "FBDynamics": ([6, 4, 4, 2], [(6, 1)]), "FBDynamics_par": ([6, 4, 4, 2], [(6, 6), (6, 4)]),
    "WB_FL1_terminal_constr": ([14], [(1, 1), (1, 14), (14, 14)]),
    "WB_FL2_terminal_constr": ([14], [(1, 1), (1, 14), (14, 14)]),
    "Jacob_F": ([14], [(2, 7), (2, 7)]), "Jacob_B": ([14], [(2, 7), (2, 7)]),
    "Link0Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link1Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link2Jacobian": ([14, 2], [(2, 7), (2, 7)]),
    "Link3Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link4Jacobian": ([14, 2], [(2, 7), (2, 7)]),
}


def eval_fn(name: str, *args, variant: str = "strict") -> List[np.ndarray]:
    """Evaluate one of the reference's generated functions; dense outputs (zeros where structurally zero)."""
    L = lib(variant)
    ins, outs = _SHAPES[name]
    a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in args]
    assert [x.size for x in a] == ins, (name, [x.size for x in a])
    o = [np.zeros(r * c) for r, c in outs]
    ap = [_dp(x) for x in a] + [None] * (4 - len(a))
    op = [_dp(x) for x in o] + [None] * (4 - len(o))
    rc = L.orc_eval(name.encode(), *ap, *op)
    assert rc == 0, name
    return [x.reshape(r, c, order="F") for x, (r, c) in zip(o, outs)]
