"""beamfoam_b200 — B200-native Newton-iteration path of beamFoam's
coupledTotalLagNewtonRaphsonBeam.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
include/beamgpu.h), the ctypes binding, and the host-side mirror of the reference's
beamModel / fvPatchField interface for this path.
"""
from . import _lib  # noqa: F401
from ._lib import BeamGpuError, load_library  # noqa: F401
from .beamModel import BeamCase, axisAngleRotation, coupledTotalLagNewtonRaphsonBeam, shard_range  # noqa: F401
from . import beamMomentumContribution, cases, crossSections, fvPatchFields  # noqa: F401

__all__ = ["BeamCase", "BeamGpuError", "axisAngleRotation", "beamMomentumContribution", "cases", "coupledTotalLagNewtonRaphsonBeam",
           "crossSections", "fvPatchFields", "load_library", "shard_range"]
