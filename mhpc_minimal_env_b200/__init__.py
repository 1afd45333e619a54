"""B200-native batched HSDDP hot path of Model-Hierarchy Predictive Control.

The package is a thin host mirror (ctypes, `BatchedMHPC`) of the C ABI in include/mhpc_b200.h; all arithmetic of the
path runs in the hand-written sm_100a CUDA kernels under csrc/ (libmhpc_b200.so).  There is no CPU fallback.
"""
from .capi import (BatchedMHPC, HSDDPOption, IterRecord, MHPCError, Problem, Result, Timing, default_option, default_problem,
                   eval_function, load, measure_fp64_peak)

__all__ = ["BatchedMHPC", "HSDDPOption", "IterRecord", "MHPCError", "Problem", "Result", "Timing", "default_option", "default_problem",
           "eval_function", "load", "measure_fp64_peak"]
