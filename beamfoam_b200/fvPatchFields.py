"""Host-side mirrors of the reference's six boundary-condition classes.

Only what stays on the host is here: dictionary-level data and the time-series
lookup of ``updateCoeffs()``.  The boundary-face arithmetic (assembly closure and
``evaluate()``) runs on the device behind the C ABI.

Reference: src/wireBunchingModels/beamModels/fvPatchFields/<name>/<name>FvPatchVectorField.C
"""
from __future__ import annotations

import bisect
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


class interpolationTable:
    """OpenFOAM ``interpolationTable<Type>`` with ``outOfBounds clamp`` (the only
    mode the tutorials use): piecewise-linear in time, clamped at both ends."""

    def __init__(self, table: Sequence):
        self.t = [float(r[0]) for r in table]
        self.v = [np.atleast_1d(np.asarray(r[1], dtype=np.float64)) for r in table]
        if any(b < a for a, b in zip(self.t, self.t[1:])):
            raise ValueError("interpolationTable: times must be non-decreasing")

    def __call__(self, t: float) -> np.ndarray:
        if t <= self.t[0]:
            return self.v[0].copy()
        if t >= self.t[-1]:
            return self.v[-1].copy()
        i = bisect.bisect_right(self.t, t) - 1
        t0, t1 = self.t[i], self.t[i + 1]
        s = (t - t0) / (t1 - t0)
        return self.v[i] + s * (self.v[i + 1] - self.v[i])


def _vec(v):
    return np.asarray(v, dtype=np.float64).reshape(3)


@dataclass
class _PatchField:
    #: optional per-beam multiplier applied to the evaluated value (synthetic bundles)
    scale: Optional[np.ndarray] = field(default=None, kw_only=True)
    code = -1
    kind = -1

    def valueAt(self, t: float) -> np.ndarray:  # what updateCoeffs() would set
        raise NotImplementedError

    @property
    def timeDependent(self) -> bool:
        """True if updateCoeffs() looks a value up in a time series; a plain ``value``/``force``/``moment`` entry is
        set once (fixedValueFvPatchField::updateCoeffs does nothing)."""
        return any(isinstance(getattr(self, k, None), interpolationTable)
                   for k in ("displacementSeries", "thetaSeries", "forceSeries", "transitionSeries",
                             "followerForceSeries", "offsetForceSeries", "momentSeries"))

    def values(self, t: float, nBeams: int) -> np.ndarray:
        v = np.broadcast_to(self.valueAt(t), (nBeams, 3)).copy()
        if self.scale is not None:
            v *= np.asarray(self.scale, dtype=np.float64).reshape(nBeams, 1)
        return v


@dataclass
class fixedValue(_PatchField):
    """OpenFOAM fixedValue (W or Theta)."""
    value: Sequence[float] = (0.0, 0.0, 0.0)
    variable: str = "W"

    @property
    def code(self):
        return L.BF_W_FIXED if self.variable == "W" else L.BF_TH_FIXED

    @property
    def kind(self):
        return L.BF_BCV_W if self.variable == "W" else L.BF_BCV_THETA

    def valueAt(self, t):
        return _vec(self.value)


@dataclass
class fixedDisplacement(_PatchField):
    """fixedDisplacementFvPatchVectorField.C:160-198 (displacementSeries)."""
    value: Sequence[float] = (0.0, 0.0, 0.0)
    displacementSeries: Optional[interpolationTable] = None
    code = L.BF_W_FIXED
    kind = L.BF_BCV_W

    def valueAt(self, t):
        return self.displacementSeries(t) if self.displacementSeries else _vec(self.value)


@dataclass
class fixedRotation(_PatchField):
    """fixedRotationFvPatchVectorField.C:160-213 (thetaSeries)."""
    value: Sequence[float] = (0.0, 0.0, 0.0)
    thetaSeries: Optional[interpolationTable] = None
    code = L.BF_TH_FIXED
    kind = L.BF_BCV_THETA

    def valueAt(self, t):
        return self.thetaSeries(t) if self.thetaSeries else _vec(self.value)


@dataclass
class forceBeamDisplacementNR(_PatchField):
    """forceBeamDisplacementNRFvPatchVectorField.C:186-209: force from forceSeries, or
    transitionSeries*force, or the constant ``force`` entry."""
    force: Sequence[float] = (0.0, 0.0, 0.0)
    forceSeries: Optional[interpolationTable] = None
    transitionSeries: Optional[interpolationTable] = None
    code = L.BF_W_FORCE
    kind = L.BF_BCV_FORCE

    def valueAt(self, t):
        if self.forceSeries:
            return self.forceSeries(t)
        if self.transitionSeries:
            return float(self.transitionSeries(t)[0]) * _vec(self.force)
        return _vec(self.force)


@dataclass
class followerForceBeamDisplacementNR(_PatchField):
    """followerForceBeamDisplacementNRFvPatchVectorField.C:183-265: the series gives the
    force in the MATERIAL frame; the device applies Lambdaf_b each Newton iteration."""
    followerForceSeries: Optional[interpolationTable] = None
    code = L.BF_W_FOLLOWER
    kind = L.BF_BCV_FORCE

    def valueAt(self, t):
        if self.followerForceSeries is None:
            raise ValueError("followerForceSeries is not defined")  # FatalError in the reference ctor
        return self.followerForceSeries(t)


@dataclass
class linearSpringForceBeamDisplacementNR(_PatchField):
    """linearSpringForceBeamDisplacementNRFvPatchVectorField.C:85-139,204-217."""
    springStiffness: float = 0.0
    springDirection: Sequence[float] = (1.0, 0.0, 0.0)
    offsetForceSeries: Optional[interpolationTable] = None
    code = L.BF_W_SPRING
    kind = L.BF_BCV_FORCE

    def valueAt(self, t):
        return self.offsetForceSeries(t) if self.offsetForceSeries else np.zeros(3)


@dataclass
class momentBeamRotationNR(_PatchField):
    """momentBeamRotationNRFvPatchVectorField.C:168-181."""
    moment: Sequence[float] = (0.0, 0.0, 0.0)
    momentSeries: Optional[interpolationTable] = None
    code = L.BF_TH_MOMENT
    kind = L.BF_BCV_MOMENT

    def valueAt(self, t):
        return self.momentSeries(t) if self.momentSeries else _vec(self.moment)
