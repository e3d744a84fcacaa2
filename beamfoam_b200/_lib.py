"""ctypes binding of include/beamgpu.h.

The product library is ``beamfoam_b200/libbeamgpu.so`` (hand-written CUDA for
sm_100a, built by ``beamfoam_b200.build``).  There is NO CPU fallback: if the
library is missing, or no CUDA device is visible when a context is created, the
call fails loudly.  Tests may pass another implementation of the same C ABI (the
CPU oracle under ``oracle/``) explicitly through ``load_library(path)``; nothing in
this package does so on its own.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(_HERE, "libbeamgpu.so")

# ---- status codes / enums (include/beamgpu.h) --------------------------------
BF_OK, BF_ERR_INVALID, BF_ERR_CUDA, BF_ERR_NOMEM = 0, 1, 2, 3
BF_DIVERGED, BF_MAXITER, BF_ERR_UNSUPPORTED, BF_ERR_COMM = 4, 5, 6, 7
BF_SCHEME_STEADY, BF_SCHEME_NEWMARK, BF_SCHEME_EULER = 0, 1, 2
BF_W_FIXED, BF_W_FORCE, BF_W_FOLLOWER, BF_W_SPRING = 0, 1, 2, 3
BF_TH_FIXED, BF_TH_MOMENT = 0, 1
BF_END_LEFT, BF_END_RIGHT = 0, 1
BF_BCV_W, BF_BCV_THETA, BF_BCV_FORCE, BF_BCV_MOMENT = 0, 1, 2, 3
FIELDS = {
    "W": 0, "Theta": 1, "Lambda": 2, "DW": 3, "DTheta": 4, "U": 5, "Accl": 6,
    "Omega": 7, "dotOmega": 8, "Lambdaf": 9, "Gamma": 10, "K": 11, "Q": 12, "M": 13,
}
FIELD_NCOMP = {"Lambda": 9, "Lambdaf": 9}
SURFACE_FIELDS = ("Lambdaf", "Gamma", "K", "Q", "M")
VOL_FIELDS_WITH_BOUNDARY = ("W", "Theta", "Lambda", "DW", "DTheta")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class bf_layout(C.Structure):
    _fields_ = [("nBeams", C.c_int64), ("nSeg", _ip), ("length", _dp), ("cellLength", _dp)]


class bf_beam_props(C.Structure):
    _fields_ = [("CQ", _dp), ("CM", _dp), ("ARho", _dp), ("CIRho", _dp), ("weight", _dp), ("refLambdaf", _dp)]


class bf_bc(C.Structure):
    _fields_ = [("wType", _ip), ("thetaType", _ip), ("springStiffness", _dp), ("springDirection", _dp)]


class bf_options(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("newmarkBeta", C.c_double), ("newmarkGamma", C.c_double),
                ("device", C.c_int32), ("storeIncrements", C.c_int32), ("reserved", C.c_int32 * 6)]


class bf_tolerances(C.Structure):
    _fields_ = [("nCorrectors", C.c_int32), ("residualTol", C.c_double), ("solutionTol", C.c_double),
                ("absoluteTol", C.c_double), ("divergenceTol", C.c_double)]


class bf_step_out(C.Structure):
    _fields_ = [("nIter", C.c_int32), ("status", C.c_int32), ("initialResidualNorm", C.c_double),
                ("currentResidualNorm", C.c_double), ("deltaXNorm", C.c_double), ("xNorm", C.c_double)]


class bf_momentum_contribution(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("coeffs", C.c_double * 6)]


BF_CONTRIB_MORISON_DRAG, BF_CONTRIB_GROUND_CONTACT, BF_MAX_CONTRIBUTIONS = 0, 1, 8

bf_reduce_fn = C.CFUNCTYPE(C.c_int, _dp, C.c_int32, C.c_void_p)

# every symbol include/beamgpu.h declares: name -> (restype, argtypes)
EXPORTS = {
    "bf_create": (C.c_int, [C.POINTER(bf_layout), C.POINTER(bf_beam_props), C.POINTER(bf_bc),
                            C.POINTER(bf_options), C.POINTER(C.c_void_p)]),
    "bf_destroy": (C.c_int, [C.c_void_p]),
    "bf_last_error": (C.c_char_p, [C.c_void_p]),
    "bf_version": (C.c_char_p, []),
    "bf_backend": (C.c_char_p, []),
    "bf_num_cells": (C.c_int64, [C.c_void_p]),
    "bf_num_internal_faces": (C.c_int64, [C.c_void_p]),
    "bf_num_beams": (C.c_int64, [C.c_void_p]),
    "bf_set_field": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_get_field": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_mirror_begin": (C.c_int, [C.c_void_p, C.c_int32, _ip, C.POINTER(C.c_void_p)]),
    "bf_mirror_wait": (C.c_int, [C.c_void_p]),
    "bf_set_distributed_loads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_set_point_forces": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_set_momentum_contributions": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(bf_momentum_contribution), C.c_void_p,
                                                C.c_void_p]),
    "bf_set_bc_values": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "bf_get_patch_pair": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "bf_set_step_outputs": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "bf_set_bc_values_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "bf_beam_energy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bf_begin_step": (C.c_int, [C.c_void_p, C.c_double]),
    "bf_newton_iteration": (C.c_int, [C.c_void_p, _dp]),
    "bf_evolve": (C.c_int, [C.c_void_p, C.POINTER(bf_tolerances), C.POINTER(bf_step_out)]),
    "bf_evolve_begin": (C.c_int, [C.c_void_p, C.POINTER(bf_tolerances)]),
    "bf_evolve_end": (C.c_int, [C.c_void_p, C.POINTER(bf_step_out)]),
    "bf_set_reducer": (C.c_int, [C.c_void_p, bf_reduce_fn, C.c_void_p]),
    "bf_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bf_comm_init_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "bf_launch_count": (C.c_int64, [C.c_void_p]),
    "bf_kernel_family": (C.c_int32, [C.c_void_p]),
    "bf_split_point": (C.c_int32, [C.c_void_p]),
    "bf_newton_iteration_checked": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_set_couplings": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bf_set_linear_solver": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double]),
    "bf_linear_iterations": (C.c_int32, [C.c_void_p]),
    "bf_comm_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bf_comm_ipc_attach": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "bf_kernel_time_ms": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, C.POINTER(C.c_int64)]),
    "bf_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "bf_region_mark": (C.c_int, [C.c_void_p, C.c_int32]),
    "bf_region_elapsed_ms": (C.c_int, [C.c_void_p, _dp]),
    "bf_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "bf_host_free": (C.c_int, [C.c_void_p]),
}

_cache: dict[str, C.CDLL] = {}


class BeamGpuError(RuntimeError):
    """Non-zero status from the C ABI (the host-side stand-in for FatalError)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"beamgpu status {status}: {message}")
        self.status = status


def load_library(path: str | None = None) -> C.CDLL:
    """Load an implementation of include/beamgpu.h and declare its prototypes.

    With ``path=None`` this is the CUDA product library; a missing library is an
    error (build it with ``python -m beamfoam_b200.build`` or ``__graft_entry__.build()``).
    """
    path = os.path.abspath(path or DEFAULT_LIBRARY)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: the CUDA extension is not built (python -m beamfoam_b200.build). "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib
