"""ctypes binding of the C ABI in include/mhpc_b200.h (the same boundary a C/C++ host binds).

The library has no CPU fallback: creating a batch without a CUDA device fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MHPC_B200_LIB") or os.path.join(HERE, "libmhpc_b200.so")   # (the override is for timing build variants)

FIELDS = ["X_ACT", "U_ACT", "Y_ACT", "X_NOM", "U_NOM", "Y_NOM", "X_REF", "U_REF", "Y_REF",
          "A", "B", "C", "D", "LX", "LU", "LY", "LXX", "LUU", "LYY", "K", "DU", "G", "H", "PHIX", "PHIXX", "PX",
          "QX", "QU", "QXX", "QUX", "QUU"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
STATUS = {0: "running", 1: "converged", 2: "max_AL_iter", 3: "regularization_abort"}


class HSDDPOption(C.Structure):
    """HSDDP_OPTION<double> (reference Common/header/MHPC_CompoundTypes.h:196-212)."""
    _fields_ = [(n, C.c_double) for n in ("alpha", "gamma", "update_penalty", "update_relax", "update_regularization",
                                          "update_ReB", "max_DDP_iter", "max_AL_iter", "DDP_thresh", "AL_thresh")] + \
               [("AL_active", C.c_int), ("ReB_active", C.c_int), ("smooth_active", C.c_int)]


class Problem(C.Structure):
    """MHPCUserParameters + USRCMD + Gait tables + x0 of one instance (include/mhpc_b200.h: mhpc_problem)."""
    _fields_ = [("n_wbphase", C.c_int), ("n_fbphase", C.c_int), ("dt_wb", C.c_float), ("dt_fb", C.c_float),
                ("cmode", C.c_int), ("groundH", C.c_float), ("vel", C.c_float), ("height", C.c_float),
                ("roll", C.c_float), ("pitch", C.c_float), ("yaw", C.c_float),
                ("gait_mode", C.c_int * 4), ("gait_timing", C.c_float * 4), ("x0", C.c_double * 14)]


class Result(C.Structure):
    _fields_ = [("cost", C.c_double), ("exp_cost_change", C.c_double), ("viol", C.c_double),
                ("status", C.c_int), ("iter_AL", C.c_int), ("n_records", C.c_int), ("pad", C.c_int)]


class IterRecord(C.Structure):
    _fields_ = [("iter_AL", C.c_int), ("iter_DDP", C.c_int), ("n_bws", C.c_int), ("n_fit", C.c_int),
                ("reg", C.c_double), ("eps", C.c_double), ("cost", C.c_double), ("exp_cost_change", C.c_double),
                ("viol", C.c_double), ("cost_prev", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("rollout_ms", C.c_float), ("linearize_ms", C.c_float), ("bws_ms", C.c_float),
                ("linesearch_ms", C.c_float), ("other_ms", C.c_float), ("n_launches", C.c_int)]


class IterTiming(C.Structure):
    _fields_ = [("iter_AL", C.c_int), ("iter_DDP", C.c_int), ("time_fwd_ms", C.c_float), ("time_bws_ms", C.c_float),
                ("time_fit_ms", C.c_float), ("time_partial_ms", C.c_float), ("total_ms", C.c_float)]


EXPORTS = [
    "mhpc_default_option", "mhpc_default_problem", "mhpc_batch_create", "mhpc_batch_destroy", "mhpc_last_error",
    "mhpc_batch_set_option", "mhpc_batch_set_stream", "mhpc_batch_set_problems", "mhpc_batch_initialization", "mhpc_batch_set_initial_condition", "mhpc_batch_update_problem",
    "mhpc_batch_forward_sweep", "mhpc_batch_forward_sweep_dynamics_only", "mhpc_batch_forward_sweep_partials_only",
    "mhpc_batch_backward_sweep", "mhpc_batch_forward_iteration", "mhpc_batch_update_nominal_trajectory",
    "mhpc_batch_update_AL_ReB_param", "mhpc_batch_solve", "mhpc_batch_solve_host", "mhpc_batch_synchronize",
    "mhpc_batch_get_results", "mhpc_batch_get_records", "mhpc_batch_get_scalars", "mhpc_batch_n_phases",
    "mhpc_batch_get_schedule", "mhpc_batch_get_field", "mhpc_batch_get_phase_params", "mhpc_batch_results_to_device",
    "mhpc_eval_function", "mhpc_batch_enable_timing", "mhpc_batch_get_timing", "mhpc_batch_time_kernel",
    "mhpc_measure_fp64_peak", "mhpc_batch_device_bytes", "mhpc_batch_enable_bws_trace", "mhpc_batch_get_bws_trace",
    "mhpc_batch_get_iteration_timing", "mhpc_batch_set_eps_smooth", "mhpc_batch_get_eps_smooth",
    "mhpc_sizeof_option", "mhpc_sizeof_problem", "mhpc_sizeof_result", "mhpc_sizeof_iter_record", "mhpc_sizeof_timing", "mhpc_struct_offsets",
]

_lib = None


def load():
    """Load libmhpc_b200.so (built by mhpc_minimal_env_b200.build). Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m mhpc_minimal_env_b200.build` "
                           "(the CUDA extension is the only implementation of the hot path)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.mhpc_last_error.restype = C.c_char_p
    L.mhpc_default_option.argtypes = [C.POINTER(HSDDPOption)]
    L.mhpc_default_problem.argtypes = [C.POINTER(Problem)]
    L.mhpc_batch_create.argtypes = [C.c_int, C.POINTER(HSDDPOption), C.c_int, C.c_int, C.POINTER(vp)]
    L.mhpc_batch_destroy.argtypes = [vp]
    L.mhpc_batch_destroy.restype = None
    L.mhpc_batch_set_option.argtypes = [vp, C.POINTER(HSDDPOption)]
    L.mhpc_batch_set_stream.argtypes = [vp, vp]
    L.mhpc_batch_set_problems.argtypes = [vp, C.POINTER(Problem)]
    L.mhpc_batch_set_initial_condition.argtypes = [vp, dp]
    for f in ("mhpc_batch_initialization", "mhpc_batch_update_problem", "mhpc_batch_forward_sweep_partials_only", "mhpc_batch_update_nominal_trajectory",
              "mhpc_batch_update_AL_ReB_param", "mhpc_batch_synchronize"):
        getattr(L, f).argtypes = [vp]
    L.mhpc_batch_forward_sweep.argtypes = [vp, C.c_double]
    L.mhpc_batch_forward_sweep_dynamics_only.argtypes = [vp, C.c_double]
    L.mhpc_batch_backward_sweep.argtypes = [vp, C.c_double, ip]
    L.mhpc_batch_forward_iteration.argtypes = [vp, ip]
    L.mhpc_batch_solve.argtypes = [vp, C.c_int]
    L.mhpc_batch_solve_host.argtypes = [vp, C.POINTER(Problem), C.POINTER(Result), C.c_int, dp, dp, C.c_int]
    L.mhpc_batch_get_results.argtypes = [vp, C.POINTER(Result)]
    L.mhpc_batch_get_records.argtypes = [vp, C.c_int, C.POINTER(IterRecord), C.c_int]
    L.mhpc_batch_get_scalars.argtypes = [vp, dp, dp, dp]
    L.mhpc_batch_n_phases.argtypes = [vp, C.c_int]
    L.mhpc_batch_get_schedule.argtypes = [vp, C.c_int, ip, ip, ip]
    L.mhpc_batch_get_field.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    L.mhpc_batch_get_phase_params.argtypes = [vp, C.c_int, C.c_int, dp]
    L.mhpc_batch_results_to_device.argtypes = [vp, vp, C.c_int]
    L.mhpc_eval_function.argtypes = [C.c_char_p, C.c_int, C.c_int] + [dp] * 8
    L.mhpc_batch_enable_bws_trace.argtypes = [vp, C.c_int]
    L.mhpc_batch_get_bws_trace.argtypes = [vp, ip]
    L.mhpc_batch_enable_timing.argtypes = [vp, C.c_int]
    L.mhpc_batch_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.mhpc_batch_set_eps_smooth.argtypes = [vp, C.c_double]
    L.mhpc_batch_get_eps_smooth.argtypes = [vp]
    L.mhpc_batch_get_eps_smooth.restype = C.c_double
    L.mhpc_batch_get_iteration_timing.argtypes = [vp, C.POINTER(IterTiming), C.c_int]
    L.mhpc_batch_time_kernel.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.mhpc_measure_fp64_peak.argtypes = [C.c_int, dp]
    L.mhpc_batch_device_bytes.argtypes = [vp]
    L.mhpc_batch_device_bytes.restype = C.c_longlong
    L.mhpc_struct_offsets.argtypes = [C.c_int, ip, C.c_int]
    # the mirrors above must have the layout the library was compiled with
    for which, (cls, szfn) in enumerate(((HSDDPOption, L.mhpc_sizeof_option), (Problem, L.mhpc_sizeof_problem), (Result, L.mhpc_sizeof_result),
                                         (IterRecord, L.mhpc_sizeof_iter_record))):
        offs = (C.c_int * 32)()
        nf = L.mhpc_struct_offsets(which, offs, 32)
        mine = [getattr(cls, f).offset for f, _ in cls._fields_]
        if C.sizeof(cls) != szfn() or nf != len(mine) or list(offs[:nf]) != mine:
            raise RuntimeError(f"ctypes mirror of {cls.__name__} does not match libmhpc_b200.so (sizeof {C.sizeof(cls)} vs {szfn()})")
    if C.sizeof(Timing) != L.mhpc_sizeof_timing():
        raise RuntimeError("ctypes mirror of Timing does not match libmhpc_b200.so")
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def default_option(**kw) -> HSDDPOption:
    o = HSDDPOption()
    load().mhpc_default_option(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_problem(**kw) -> Problem:
    p = Problem()
    load().mhpc_default_problem(C.byref(p))
    for k, v in kw.items():
        if k in ("gait_mode", "gait_timing", "x0"):
            getattr(p, k)[:] = list(v)
        else:
            setattr(p, k, v)
    return p


class MHPCError(RuntimeError):
    pass


class BatchedMHPC:
    """Batched counterpart of the reference's MHPCLocomotion<double> / MultiPhaseDDP<double>
    (Controller/MHPCLocomotion/MHPCLocomotion.h:13-83, HSDDPSolver/header/MultiPhaseDDP.h:8-69): same method
    names and meaning, every call acting on all instances of the batch."""

    def __init__(self, problems: Sequence[Problem], option: Optional[HSDDPOption] = None, device: int = 0, ls_width: int = 8):
        self.L = load()
        self.n = len(problems)
        self.option = option if option is not None else default_option()
        self.h = C.c_void_p()
        self._ck(self.L.mhpc_batch_create(self.n, C.byref(self.option), device, ls_width, C.byref(self.h)))
        self._problems = (Problem * self.n)(*problems)
        self._ck(self.L.mhpc_batch_set_problems(self.h, self._problems))

    def _ck(self, rc):
        if rc < 0:
            raise MHPCError(self.L.mhpc_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.mhpc_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- MHPCLocomotion / MultiPhaseDDP surface ----
    def set_option(self, option: HSDDPOption):
        self.option = option
        self._ck(self.L.mhpc_batch_set_option(self.h, C.byref(option)))

    def set_problems(self, problems: Sequence[Problem]):
        assert len(problems) == self.n
        self._problems = (Problem * self.n)(*problems)
        self._ck(self.L.mhpc_batch_set_problems(self.h, self._problems))

    def set_stream(self, cuda_stream_ptr: int): self._ck(self.L.mhpc_batch_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))
    def initialization(self): self._ck(self.L.mhpc_batch_initialization(self.h))
    def update_problem(self): self._ck(self.L.mhpc_batch_update_problem(self.h))

    def set_initial_condition(self, x0):
        a = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(self.n, 14))
        self._ck(self.L.mhpc_batch_set_initial_condition(self.h, _dp(a)))
    def forward_sweep(self, eps: float): self._ck(self.L.mhpc_batch_forward_sweep(self.h, eps))
    def forward_sweep_dynamics_only(self, eps: float): self._ck(self.L.mhpc_batch_forward_sweep_dynamics_only(self.h, eps))
    def forward_sweep_partials_only(self): self._ck(self.L.mhpc_batch_forward_sweep_partials_only(self.h))

    def backward_sweep(self, regularization: float) -> np.ndarray:
        ok = np.zeros(self.n, dtype=np.int32)
        self._ck(self.L.mhpc_batch_backward_sweep(self.h, regularization, ok.ctypes.data_as(C.POINTER(C.c_int))))
        return ok.astype(bool)

    def forward_iteration(self) -> np.ndarray:
        t = np.zeros(self.n, dtype=np.int32)
        self._ck(self.L.mhpc_batch_forward_iteration(self.h, t.ctypes.data_as(C.POINTER(C.c_int))))
        return t

    def update_nominal_trajectory(self): self._ck(self.L.mhpc_batch_update_nominal_trajectory(self.h))
    def update_AL_ReB_param(self): self._ck(self.L.mhpc_batch_update_AL_ReB_param(self.h))
    def solve(self, sync_early_exit: bool = False): self._ck(self.L.mhpc_batch_solve(self.h, int(sync_early_exit)))
    def synchronize(self): self._ck(self.L.mhpc_batch_synchronize(self.h))

    def solve_host(self, problems=None, exec_knots: int = 0):
        """solve_mhpc end to end from host buffers; returns (results, exec_state, exec_gain)."""
        pr = self._problems if problems is None else (Problem * self.n)(*problems)
        res = (Result * self.n)()
        st = np.zeros((self.n, exec_knots, 22)) if exec_knots else None
        gn = np.zeros((self.n, exec_knots, 60)) if exec_knots else None
        self._ck(self.L.mhpc_batch_solve_host(self.h, pr, res, 2 if exec_knots else 0, _dp(st), _dp(gn), exec_knots))
        return res, st, gn

    # ---- results ----
    def results(self) -> List[dict]:
        res = (Result * self.n)()
        self._ck(self.L.mhpc_batch_get_results(self.h, res))
        return [{"cost": r.cost, "exp_cost_change": r.exp_cost_change, "viol": r.viol, "status": r.status,
                 "iter_AL": r.iter_AL, "n_records": r.n_records} for r in res]

    def results_array(self):
        res = (Result * self.n)()
        self._ck(self.L.mhpc_batch_get_results(self.h, res))
        return np.frombuffer(res, dtype=np.dtype([("cost", "f8"), ("exp_cost_change", "f8"), ("viol", "f8"),
                                                  ("status", "i4"), ("iter_AL", "i4"), ("n_records", "i4"), ("pad", "i4")])).copy()

    def scalars(self):
        c, d, v = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n)
        self._ck(self.L.mhpc_batch_get_scalars(self.h, _dp(c), _dp(d), _dp(v)))
        return {"cost": c, "exp_cost_change": d, "viol": v}

    def records(self, instance: int) -> List[dict]:
        buf = (IterRecord * 256)()
        n = self._ck(self.L.mhpc_batch_get_records(self.h, instance, buf, 256))
        return [{"iter_AL": r.iter_AL, "iter_DDP": r.iter_DDP, "n_bws": r.n_bws, "n_fit": r.n_fit, "reg": r.reg, "eps": r.eps,
                 "cost": r.cost, "exp_cost_change": r.exp_cost_change, "viol": r.viol, "cost_prev": r.cost_prev} for r in buf[:n]]

    def schedule(self, instance: int = 0):
        n = self._ck(self.L.mhpc_batch_n_phases(self.h, instance))
        m, N, w = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
        self._ck(self.L.mhpc_batch_get_schedule(self.h, instance, m, N, w))
        return list(m), list(N), [bool(x) for x in w]

    def field(self, instance: int, phase: int, name: str) -> np.ndarray:
        fid = FIELD_ID[name]
        dim = self._ck(self.L.mhpc_batch_get_field(self.h, instance, phase, fid, None))
        _, N, _ = self.schedule(instance)
        out = np.zeros(dim) if name in ("PHIX", "PHIXX", "PX") else np.zeros((N[phase], dim))
        self._ck(self.L.mhpc_batch_get_field(self.h, instance, phase, fid, _dp(out)))
        return out

    def phase_params(self, instance: int, phase: int):
        o = np.zeros(9)
        self._ck(self.L.mhpc_batch_get_phase_params(self.h, instance, phase, _dp(o)))
        return {"sigma": o[0], "lambda": o[1], "delta": list(o[2:5]), "eps": list(o[5:8]), "h": o[8]}

    def results_to_device(self, dev_ptr: int, capacity_doubles: int) -> int:
        return self._ck(self.L.mhpc_batch_results_to_device(self.h, C.c_void_p(dev_ptr), capacity_doubles))

    # ---- measurement ----
    def enable_timing(self, on: bool = True): self._ck(self.L.mhpc_batch_enable_timing(self.h, int(on)))

    def enable_bws_trace(self, on: bool = True): self._ck(self.L.mhpc_batch_enable_bws_trace(self.h, int(on)))

    def bws_trace(self, max_rec: int = 64, att: int = 24) -> np.ndarray:
        """[n, record, attempt]: backward step + 1 at which the attempt failed, -1 passed, 0 not made."""
        out = np.zeros((self.n, max_rec * att), dtype=np.int32)
        rows = self._ck(self.L.mhpc_batch_get_bws_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_int))))
        assert rows == max_rec * att, rows
        return out.reshape(self.n, max_rec, att)

    def timing(self) -> dict:
        t = Timing()
        self._ck(self.L.mhpc_batch_get_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def set_eps_smooth(self, v: float): self._ck(self.L.mhpc_batch_set_eps_smooth(self.h, float(v)))
    def eps_smooth(self) -> float: return float(self.L.mhpc_batch_get_eps_smooth(self.h))

    def iteration_timing(self) -> List[dict]:
        buf = (IterTiming * 256)()
        n = self._ck(self.L.mhpc_batch_get_iteration_timing(self.h, buf, 256))
        return [{k: getattr(r, k) for k, _ in IterTiming._fields_} for r in buf[:n]]

    def time_kernel(self, which: int, reps: int) -> float:
        ms = C.c_float()
        self._ck(self.L.mhpc_batch_time_kernel(self.h, which, reps, C.byref(ms)))
        return ms.value

    def device_bytes(self) -> int: return int(self.L.mhpc_batch_device_bytes(self.h))


_EVAL_SHAPES = {
    "Dyn_BS": ([14, 4], [(14, 1), (4, 1)]), "Dyn_FS": ([14, 4], [(14, 1), (4, 1)]), "Dyn_FL": ([14, 4], [(14, 1), (4, 1)]),
    "Imp_F": ([14], [(14, 1), (4, 1)]), "Imp_B": ([14], [(14, 1), (4, 1)]),
    "Dyn_BS_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]), "Dyn_FS_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]),
    "Dyn_FL_par": ([14, 4], [(14, 14), (14, 4), (4, 14), (4, 4)]),
    "Imp_F_par": ([14], [(14, 14)]), "Imp_B_par": ([14], [(14, 14)]),
    "FBDynamics": ([6, 4, 4, 2], [(6, 1)]), "FBDynamics_par": ([6, 4, 4, 2], [(6, 6), (6, 4)]),
    "WB_FL1_terminal_constr": ([14], [(1, 1), (1, 14), (14, 14)]), "WB_FL2_terminal_constr": ([14], [(1, 1), (1, 14), (14, 14)]),
    "Jacob_F": ([14], [(2, 7)]), "Jacob_B": ([14], [(2, 7)]),
    "Link0Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link1Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link2Jacobian": ([14, 2], [(2, 7), (2, 7)]),
    "Link3Jacobian": ([14, 2], [(2, 7), (2, 7)]), "Link4Jacobian": ([14, 2], [(2, 7), (2, 7)]),
}


def eval_function(name: str, *args, device: int = 0):
    """Evaluate one of the generated functions (CasADi C ABI of the reference, CasadiGen/header/*.h) for a batch of
    inputs ([n, size] arrays) on the device; dense outputs [n, rows, cols]."""
    L = load()
    ins, outs = _EVAL_SHAPES[name]
    a = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in args]
    n = a[0].shape[0]
    assert [x.shape[1] for x in a] == ins
    o = [np.zeros((n, r * c)) for r, c in outs]
    ap = [_dp(x) for x in a] + [None] * (4 - len(a))
    op = [_dp(x) for x in o] + [None] * (4 - len(o))
    rc = L.mhpc_eval_function(name.encode(), n, device, *ap, *op)
    if rc < 0:
        raise MHPCError(L.mhpc_last_error().decode())
    return [x.reshape(n, c, r).transpose(0, 2, 1) for x, (r, c) in zip(o, outs)]


def measure_fp64_peak(device: int = 0) -> float:
    t = C.c_double()
    L = load()
    if L.mhpc_measure_fp64_peak(device, C.byref(t)) < 0:
        raise MHPCError(L.mhpc_last_error().decode())
    return t.value
