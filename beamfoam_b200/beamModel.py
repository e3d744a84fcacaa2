"""Host-side mirror of the reference's ``beamModel`` /
``coupledTotalLagNewtonRaphsonBeam`` interface for the Newton-iteration path.

The class keeps the reference's names and call order
(``beamModel::New`` → ``evolve()`` per time step → ``solutionW()`` ...;
applications/solvers/beamFoam/beamFoam.C:57-94) but owns no numerics: every
per-cell / per-face operation happens behind the C ABI of include/beamgpu.h in
``libbeamgpu.so`` (CUDA, sm_100a).  The host part is what stays on the host in the
reference as well: dictionary-level set-up (constant/beamProperties, 0/W, 0/Theta,
system/fvSchemes), the BC time series, and the time loop.

OpenFOAM itself is not available in this environment, so the case description is
a plain-Python ``BeamCase`` instead of dictionaries on disk; INTEGRATION.md shows
the OpenFOAM-side subclass that feeds the same C ABI from real dictionaries.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence, Union

import numpy as np

from . import _lib as L
from ._lib import BeamGpuError
from .fvPatchFields import _PatchField, fixedValue


def axisAngleRotation(axis: Sequence[float], angleDeg: float) -> np.ndarray:
    """``coordinateRotations::axisAngle::rotation(axis, angle, degrees=true)`` as used by
    setInitialPositionBeam (applications/utilities/setInitialPositionBeam/setInitialPositionBeam.C:197-212):
    Rodrigues rotation about the normalised axis."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    th = math.radians(angleDeg)
    S = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(th) * S + (1 - math.cos(th)) * (S @ S)


@dataclass
class BeamCase:
    """Everything the reference reads from a case directory for this path.

    Per-beam quantities are arrays of length nBeams (a scalar is broadcast), so a
    65,536-beam bundle costs a handful of numpy arrays, not 65,536 objects.
    """
    # constant/beamProperties: beams ( beam_k { crossSectionModel ...; length; nSegments; E; G; rho } )
    nSegments: Union[int, Sequence[int]]
    length: Union[float, Sequence[float]]
    A: Union[float, Sequence[float]]      # crossSection.A()
    Iyy: Union[float, Sequence[float]]    # beamModel::Iyy = crossSection.Ixx()  (beamModel.H:449-461)
    Izz: Union[float, Sequence[float]]    # beamModel::Izz = crossSection.Iyy()  (beamModel.H:463-475)
    J: Union[float, Sequence[float]]      # crossSection.IT()
    E: Union[float, Sequence[float]]
    G: Union[float, Sequence[float]]
    rho: Union[float, Sequence[float]] = 0.0
    nBeams: int = 1
    # boundary conditions: one patch-field object per end (applies to every beam) or a list per beam
    W_left: Union[_PatchField, Sequence[_PatchField]] = field(default_factory=lambda: fixedValue(variable="W"))
    W_right: Union[_PatchField, Sequence[_PatchField]] = field(default_factory=lambda: fixedValue(variable="W"))
    Theta_left: Union[_PatchField, Sequence[_PatchField]] = field(default_factory=lambda: fixedValue(variable="Theta"))
    Theta_right: Union[_PatchField, Sequence[_PatchField]] = field(default_factory=lambda: fixedValue(variable="Theta"))
    # system/fvSchemes d2dt2Schemes
    d2dt2Scheme: str = "steadyState"
    newmarkBeta: float = 0.25
    newmarkGamma: float = 0.5
    # <model>Coeffs
    nCorrectors: int = 1000
    residualTol: float = -1.0
    convergenceTol: Optional[float] = None
    solutionTol: float = 1e-10
    absoluteTol: float = 1e-30
    divergenceTol: float = 1e6
    scalingInertiaTensor: float = 1.0
    # constant/g, 0/q, 0/m
    g: Sequence[float] = (0.0, 0.0, 0.0)
    q: Optional[Sequence[float]] = None   # uniform (3,) or per-cell (nCells,3)
    m: Optional[Sequence[float]] = None
    # constant/pointForces (beamModel.C:863-866): ((cellI zeta) force) with cellI the cell label in the bundle,
    # zeta in [-1, 1] and force a 3-vector or an interpolationTable
    pointForces: Optional[Sequence[tuple]] = None
    # setInitialPositionBeam -rotateAngle '(axis angle)': 3x3 T (all beams) or (nBeams,3,3); None = I
    initialRotation: Optional[np.ndarray] = None
    # setInitialPositionBeam -translate '(x y z)': (3,) or (nBeams,3); only the momentum contributions read it
    initialTranslation: Optional[np.ndarray] = None
    # constant/beamMomentumContributionProperties: list of morisonDragContribution / groundContactContribution
    momentumContributions: Optional[Sequence] = None
    # inter-beam couplings (synthetic, see include/beamgpu.h bf_set_couplings): (cellA, cellB, stiffness)
    # with cell labels of the whole bundle; linearSolver = (maxIter, tolerance, relTol) of the coupled BiCGStab
    couplings: Optional[Sequence[tuple]] = None
    linearSolver: Optional[tuple] = None
    # crossSection.R() per beam (circle: radius, rectangle: h/2): read by the momentum contributions only
    R: Union[float, Sequence[float]] = 0.0
    # system/controlDict
    deltaT: float = 1.0
    endTime: float = 1.0
    name: str = "case"

    def per_beam(self, x, dtype=np.float64) -> np.ndarray:
        return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=dtype), (self.nBeams,)))

    def rtol(self) -> float:
        # Evolve.C:67-79: residualTol, else convergenceTol, else 1e-8
        if self.residualTol >= 0:
            return self.residualTol
        return self.convergenceTol if self.convergenceTol is not None else 1e-8


def shard_range(nBeams: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous beam range of ``rank``: beam b lives on rank floor(b*world/nBeams)
    (SURVEY §8e)."""
    lo = -(-rank * nBeams // world)
    hi = -(-(rank + 1) * nBeams // world)
    return lo, min(hi, nBeams)


class coupledTotalLagNewtonRaphsonBeam:
    """Drop-in host object for the reference class of the same name.

    ``library`` selects the implementation of include/beamgpu.h; the default is the
    CUDA product library and nothing else is ever chosen automatically.
    ``beamRange=(lo, hi)`` makes this object own only that shard of the bundle;
    ``reducer(np.ndarray[3]) -> None`` then sums the three squared norms over all
    shards in place (NCCL all-reduce in bench.py, gloo in the CPU tests).
    """

    typeName = "coupledTotalLagNewtonRaphsonBeam"

    def __init__(self, case: BeamCase, library: Optional[C.CDLL] = None, device: int = 0,
                 beamRange: Optional[tuple[int, int]] = None,
                 reducer: Optional[Callable[[np.ndarray], None]] = None,
                 storeIncrements: bool = True):
        self.case = case
        self.lib = library if library is not None else L.load_library()
        lo, hi = beamRange if beamRange is not None else (0, case.nBeams)
        if not (0 <= lo < hi <= case.nBeams):
            raise ValueError(f"empty or invalid beam range {(lo, hi)} for {case.nBeams} beams")
        self.beamRange = (lo, hi)
        B = self._B = hi - lo
        sl = slice(lo, hi)

        scheme = {"steadyState": L.BF_SCHEME_STEADY, "Newmark": L.BF_SCHEME_NEWMARK,
                  "Euler": L.BF_SCHEME_EULER}.get(case.d2dt2Scheme)
        if scheme is None:  # CTL.C:708-720
            raise ValueError("Please specify the d2dt2Schemes: steadyState, Euler, Newmark")

        nSeg = case.per_beam(case.nSegments, np.int32)[sl].copy()
        length = case.per_beam(case.length)[sl].copy()
        A, Iyy, Izz, J = (case.per_beam(x)[sl] for x in (case.A, case.Iyy, case.Izz, case.J))
        E, G, rho = (case.per_beam(x)[sl] for x in (case.E, case.G, case.rho))
        if scheme != L.BF_SCHEME_STEADY and np.any(rho == 0.0):  # Evolve.C:323-335
            raise ValueError(f"The time integration scheme provided is {case.d2dt2Scheme} "
                             "but density 'rho' is not specified!!")
        # CTL.C:538-585 (CQ_, CM_), :621-667 (ARho_, CIRho_), Evolve.C:201-209 (self-weight)
        CQ = np.ascontiguousarray(np.stack([E * A, G * A, G * A], axis=1))
        CM = np.ascontiguousarray(np.stack([G * J, E * Iyy, E * Izz], axis=1))
        ARho = np.ascontiguousarray(A * rho)
        kCI = case.scalingInertiaTensor
        CIRho = np.ascontiguousarray(np.stack([kCI * (rho * J), kCI * (rho * Iyy), kCI * (rho * Izz)], axis=1))
        g = np.asarray(case.g, dtype=np.float64).reshape(3)
        weight = np.ascontiguousarray((rho * A)[:, None] * g[None, :])
        refL = None
        if case.initialRotation is not None:
            T = np.asarray(case.initialRotation, dtype=np.float64)
            T = np.broadcast_to(T, (case.nBeams, 3, 3))[sl] if T.ndim == 2 else T[sl]
            refL = np.ascontiguousarray(T.reshape(B, 9))

        self._bcs = {}
        self._bcCache = {}
        self._bcUploaded = set()
        self._pinnedBC = {}
        self._bcSet = 0        # which of the two sets of staging arrays updateCoeffs() fills
        self._bcPrepared = []  # (end, kind, staging array) evaluated ahead by updateCoeffs(t, prepareOnly=True)
        self._pinnedPtrs = []
        self.h2dBytes = 0  # bytes handed to bf_set_bc_values so far
        self._beganStepAt = None  # time index of the last bf_begin_step
        wType = np.zeros(2 * B, np.int32)
        thType = np.zeros(2 * B, np.int32)
        springK = np.zeros(2 * B)
        springDir = np.zeros((2 * B, 3))
        for end, (wspec, tspec) in enumerate(((case.W_left, case.Theta_left), (case.W_right, case.Theta_right))):
            for var, spec, types in (("W", wspec, wType), ("Theta", tspec, thType)):
                objs = spec if isinstance(spec, (list, tuple)) else None
                if objs is not None:
                    if len(objs) != case.nBeams:
                        raise ValueError("per-beam patch-field list must have nBeams entries")
                    objs = list(objs[lo:hi])
                    types[end * B:(end + 1) * B] = [o.code for o in objs]
                else:
                    types[end * B:(end + 1) * B] = spec.code
                self._bcs[(end, var)] = (spec, objs)
                if var == "W":
                    for b in range(B):
                        o = objs[b] if objs is not None else spec
                        if o.code == L.BF_W_SPRING:
                            springK[end * B + b] = o.springStiffness
                            springDir[end * B + b] = o.springDirection
                    if objs is None and spec.code != L.BF_W_SPRING:
                        pass

        self._keep = (nSeg, length, CQ, CM, ARho, CIRho, weight, refL, wType, thType, springK, springDir)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None  # noqa: E731
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
        lay = L.bf_layout(B, ip(nSeg), dp(length), None)
        props = L.bf_beam_props(dp(CQ), dp(CM), dp(ARho), dp(CIRho), dp(weight), dp(refL))
        bc = L.bf_bc(ip(wType), ip(thType), dp(springK), dp(springDir))
        opt = L.bf_options(scheme, case.newmarkBeta, case.newmarkGamma, device, int(storeIncrements))
        ctx = C.c_void_p()
        rc = self.lib.bf_create(C.byref(lay), C.byref(props), C.byref(bc), C.byref(opt), C.byref(ctx))
        if rc != L.BF_OK:
            raise BeamGpuError(rc, self.lib.bf_last_error(None).decode())
        self._ctx = ctx
        self.nSeg = nSeg
        self.cellStart = np.concatenate([[0], np.cumsum(nSeg)]).astype(np.int64)
        self.nCells = int(self.lib.bf_num_cells(ctx))
        self.nInternalFaces = int(self.lib.bf_num_internal_faces(ctx))

        # patch values read from 0/W, 0/Theta ("value uniform ...") for fixedValue patches
        for var in ("W", "Theta"):
            left = self._bc_values(0, var, 0.0, only_fixed=True)
            right = self._bc_values(1, var, 0.0, only_fixed=True)
            if np.any(left) or np.any(right):
                self.set_field(var, None, left, right)
        if case.q is not None or case.m is not None:
            self._set_loads(case.q, case.m, lo, hi)
        if case.couplings:
            allStart = np.concatenate([[0], np.cumsum(case.per_beam(case.nSegments, np.int64))])
            c0, c1 = int(allStart[lo]), int(allStart[hi])
            cp = np.asarray([(a, b, k) for a, b, k in case.couplings], dtype=np.float64)
            a, b = cp[:, 0].astype(np.int64), cp[:, 1].astype(np.int64)
            inside = (a >= c0) & (a < c1) & (b >= c0) & (b < c1)
            if np.any(((a >= c0) & (a < c1)) != ((b >= c0) & (b < c1))):
                raise ValueError("a coupling pair spans two shards: coupled beams must live in one context")
            a, b, k = np.ascontiguousarray(a[inside] - c0), np.ascontiguousarray(b[inside] - c0), np.ascontiguousarray(cp[inside, 2])
            self._check(self.lib.bf_set_couplings(ctx, len(a), a.ctypes.data, b.ctypes.data, k.ctypes.data))
        if case.linearSolver:
            self._check(self.lib.bf_set_linear_solver(ctx, int(case.linearSolver[0]), float(case.linearSolver[1]), float(case.linearSolver[2])))

        if case.momentumContributions:
            lst = (L.bf_momentum_contribution * len(case.momentumContributions))()
            for k, mc in enumerate(case.momentumContributions):
                lst[k].type = mc.type
                for j, v in enumerate(mc.coeffs()):
                    lst[k].coeffs[j] = float(v)
            radius = np.ascontiguousarray(case.per_beam(case.R)[sl])
            tr = None
            if case.initialTranslation is not None:
                tr = np.asarray(case.initialTranslation, dtype=np.float64)
                tr = np.ascontiguousarray(np.broadcast_to(tr, (case.nBeams, 3))[sl] if tr.ndim == 1 else tr[sl])
            self._keep += (lst, radius, tr)
            self._check(self.lib.bf_set_momentum_contributions(
                ctx, len(case.momentumContributions), lst, radius.ctypes.data, tr.ctypes.data if tr is not None else None))

        self.tol = L.bf_tolerances(case.nCorrectors, case.rtol(), case.solutionTol, case.absoluteTol,
                                   case.divergenceTol)
        self._reducer_cb = None
        if reducer is not None:
            def _cb(ptr, n, _user, _r=reducer):
                try:
                    arr = np.ctypeslib.as_array(ptr, shape=(n,))
                    _r(arr)
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    return 1
            self._reducer_cb = L.bf_reduce_fn(_cb)
            self._check(self.lib.bf_set_reducer(ctx, self._reducer_cb, None))

        self.time = 0.0
        self.timeIndex = 0
        self._iOuterCorr = 0
        self.totalIter_ = 0
        self.lastStep: Optional[L.bf_step_out] = None

    # ------------------------------------------------------------------ utils
    def _check(self, rc: int):
        if rc != L.BF_OK:
            raise BeamGpuError(rc, self.lib.bf_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.bf_destroy(self._ctx)
            self._ctx = None
            self._pinnedBC = {}
            for ptr in self._pinnedPtrs:
                self.lib.bf_host_free(ptr)
            self._pinnedPtrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _bc_values(self, end: int, var: str, t: float, only_fixed: bool = False) -> np.ndarray:
        spec, objs = self._bcs[(end, var)]
        lo, hi = self.beamRange
        B = self._B
        fixed_code = L.BF_W_FIXED if var == "W" else L.BF_TH_FIXED
        if objs is None:
            if only_fixed and spec.code != fixed_code:
                return np.zeros((B, 3))
            # the expansion to [nBeams, 3] is cached on the evaluated 3-vector: a series that did not
            # move since the last step costs nothing on the host (the values are still uploaded)
            base = spec.valueAt(t)
            key = (end, var)
            hit = self._bcCache.get(key)
            if hit is not None and np.array_equal(hit[0], base):
                return hit[1]
            v = np.ascontiguousarray(spec.values(t, self.case.nBeams)[lo:hi])
            self._bcCache[key] = (base.copy(), v)
            return v
        out = np.zeros((B, 3))
        for b, o in enumerate(objs):
            if only_fixed and o.code != fixed_code:
                continue
            out[b] = o.valueAt(t)
        return out

    def _set_loads(self, q, m, lo, hi):
        def expand(x):
            if x is None:
                return None
            x = np.asarray(x, dtype=np.float64)
            if x.ndim == 1:
                return np.ascontiguousarray(np.broadcast_to(x, (self.nCells, 3)))
            allStart = np.concatenate([[0], np.cumsum(self.case.per_beam(self.case.nSegments, np.int64))])
            return np.ascontiguousarray(x[allStart[lo]:allStart[hi]])
        qa, ma = expand(q), expand(m)
        self._check(self.lib.bf_set_distributed_loads(
            self._ctx, qa.ctypes.data if qa is not None else None, ma.ctypes.data if ma is not None else None))

    # ------------------------------------------------- field mirroring (host)
    def _ncomp(self, name):
        return L.FIELD_NCOMP.get(name, 3)

    def get_field(self, name: str, internal: bool = True, boundary: bool = True):
        """(internal, left, right) numpy arrays in the reference's ordering."""
        nc = self._ncomp(name)
        n_int = self.nInternalFaces if name in L.SURFACE_FIELDS else self.nCells
        has_b = boundary and (name in L.SURFACE_FIELDS or name in L.VOL_FIELDS_WITH_BOUNDARY)
        I = np.empty((n_int, nc)) if internal else None
        Lf = np.empty((self._B, nc)) if has_b else None
        R = np.empty((self._B, nc)) if has_b else None
        self._check(self.lib.bf_get_field(
            self._ctx, L.FIELDS[name], I.ctypes.data if I is not None else None,
            Lf.ctypes.data if Lf is not None else None, R.ctypes.data if R is not None else None))
        return I, Lf, R

    def set_field(self, name: str, internal=None, left=None, right=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (internal, left, right)]
        self._check(self.lib.bf_set_field(self._ctx, L.FIELDS[name],
                                          *[a.ctypes.data if a is not None else None for a in arrs]))

    # reference accessors (beamModel.H:637-712)
    def solutionW(self):
        return self.get_field("W")

    def solutionDW(self):
        return self.get_field("DW")

    def solutionTheta(self):
        return self.get_field("Theta")

    def solutionLambda(self):
        return self.get_field("Lambdaf")

    def solutionRMC(self):
        return self.get_field("Lambda")

    def iOuterCorr(self) -> int:
        return self._iOuterCorr

    def beamEnergy(self) -> np.ndarray:
        """functionObjects/beamEnergyData (beamEnergyData.C:84-215): [nBeams, 3] = internal (strain) energy,
        linear kinetic energy, angular kinetic energy of every beam of this shard."""
        out = np.zeros((self._B, 3))
        self._check(self.lib.bf_beam_energy(self._ctx, out.ctypes.data))
        return out

    def tip(self, end: int = 1):
        """Patch values (W_b, Theta_b) of every beam at one end — the parity observable."""
        w, t = np.empty((self._B, 3)), np.empty((self._B, 3))
        self._check(self.lib.bf_get_patch_pair(self._ctx, end, w.ctypes.data, t.ctypes.data))
        return w, t

    # ------------------------------------------------------------- hot path
    def updateCoeffs(self, t: float, prepareOnly: bool = False):
        """W_/Theta_.boundaryFieldRef().updateCoeffs() (Evolve.C:156-157): evaluate every
        BC series at time t and hand the values to the device.  Patch fields without a series are handed over once
        (their updateCoeffs() does nothing in the reference either).
        prepareOnly: evaluate into the OTHER set of page-locked staging arrays and keep the hand-over for uploadPrepared() --
        what a host does between evolveBegin() and evolveEnd(), while the GPU runs the Newton loop of the open step."""
        if prepareOnly:
            self._bcSet ^= 1
        for end in (0, 1):
            for var in ("W", "Theta"):
                spec, objs = self._bcs[(end, var)]
                dependent = spec.timeDependent if objs is None else any(o.timeDependent for o in objs)
                if not dependent and (end, var) in self._bcUploaded:
                    continue
                vals = self._bc_values(end, var, t)
                kinds = {spec.kind} if objs is None else {o.kind for o in objs}
                for kind in kinds:
                    v = vals
                    if objs is not None and len(kinds) > 1:
                        v = np.where(np.array([o.kind == kind for o in objs])[:, None], vals, 0.0)
                    buf = self._pinned_bc(end, var, kind)  # page-locked staging: the H2D copy is a plain DMA
                    np.copyto(buf, v)
                    if prepareOnly:
                        self._bcPrepared.append((end, kind, buf))
                        continue
                    # the staging array is page-locked and refilled once per step: the upload need not be waited for
                    self._check(self.lib.bf_set_bc_values_async(self._ctx, end, kind, buf.ctypes.data))
                    self.h2dBytes += buf.nbytes
                self._bcUploaded.add((end, var))

    def uploadPrepared(self):
        """Hands the values of updateCoeffs(t, prepareOnly=True) to the device (after evolveEnd())."""
        for end, kind, buf in self._bcPrepared:
            self._check(self.lib.bf_set_bc_values_async(self._ctx, end, kind, buf.ctypes.data))
            self.h2dBytes += buf.nbytes
        self._bcPrepared = []

    def stepOutputs(self, end: int = 1):
        """Registers (once) and returns two page-locked [nBeams, 3] arrays that every evolve() fills with the patch values of W
        and Theta of `end` before it returns (bf_set_step_outputs): the per-step read of functionObjects/beamDisplacements
        without a second synchronisation."""
        if getattr(self, "_stepOut", None) is None or self._stepOut[0] != end:
            W, T = self.pinned((self._B, 3)), self.pinned((self._B, 3))
            self._check(self.lib.bf_set_step_outputs(self._ctx, end, W.ctypes.data, T.ctypes.data))
            self._stepOut = (end, W, T)
        return self._stepOut[1], self._stepOut[2]

    def pinned(self, shape) -> np.ndarray:
        """A page-locked float64 host array owned by this object (bf_host_alloc), e.g. for bf_get_patch_pair."""
        n = int(np.prod(shape))
        ptr = C.c_void_p()
        self._check(self.lib.bf_host_alloc(C.byref(ptr), n * 8))
        self._pinnedPtrs.append(ptr)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=tuple(shape))

    def _pinned_bc(self, end: int, var: str, kind: int) -> np.ndarray:
        """[nBeams, 3] page-locked host array (bf_host_alloc: cudaHostAlloc in libbeamgpu, malloc in the oracle) that the
        boundary values of one patch side are staged in before they cross the ABI."""
        key = (end, var, kind, self._bcSet)
        buf = self._pinnedBC.get(key)
        if buf is None:
            ptr = C.c_void_p()
            self._check(self.lib.bf_host_alloc(C.byref(ptr), self._B * 3 * 8))
            self._pinnedPtrs.append(ptr)
            buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(self._B, 3))
            self._pinnedBC[key] = buf
        return buf

    def updatePointForces(self, t: float):
        """pointForces()[pfI].second()(runTime().value()) (Evolve.C:219): evaluate the point-force series at
        time t for the cells of this shard and hand them to the device."""
        pfs = self.case.pointForces
        if not pfs:
            return
        lo, hi = self.beamRange
        allStart = np.concatenate([[0], np.cumsum(self.case.per_beam(self.case.nSegments, np.int64))])
        c0, c1 = int(allStart[lo]), int(allStart[hi])
        cells, zetas, forces = [], [], []
        for cell, zeta, f in pfs:
            if c0 <= cell < c1:
                cells.append(cell - c0)
                zetas.append(float(zeta))
                forces.append(np.asarray(f.valueAt(t) if hasattr(f, "valueAt") else (f(t) if callable(f) else f), dtype=np.float64))
        ca = np.ascontiguousarray(cells, dtype=np.int64)
        za = np.ascontiguousarray(zetas, dtype=np.float64)
        fa = np.ascontiguousarray(np.reshape(forces, (-1, 3)) if forces else np.zeros((0, 3)), dtype=np.float64)
        self._check(self.lib.bf_set_point_forces(self._ctx, len(cells), ca.ctypes.data, za.ctypes.data, fa.ctypes.data))

    def evolve(self) -> float:
        """``scalar evolve()`` (Evolve.C:55-670).  Returns the initial residual norm; a
        diverged / non-converged step raises (FatalError in the reference)."""
        self.updateCoeffs(self.time)
        self.updatePointForces(self.time)
        if self._beganStepAt != self.timeIndex:
            # runTime++ happened since the last evolve(): new old-time level, Newmark predictor on the first iteration.
            # A second evolve() at the same time index (FSI outer iterations) continues from the current state.
            self._check(self.lib.bf_begin_step(self._ctx, float(self.case.deltaT)))
            self._beganStepAt = self.timeIndex
        out = L.bf_step_out()
        rc = self.lib.bf_evolve(self._ctx, C.byref(self.tol), C.byref(out))
        self.lastStep = out
        self._iOuterCorr = out.nIter
        self.totalIter_ += out.nIter
        self._check(rc)
        return out.initialResidualNorm

    def evolveBegin(self):
        """First half of evolve() for a host that prepares the next time level while the GPU works (bf_evolve_begin): the
        boundary values of THIS time level must have been handed over (updateCoeffs / uploadPrepared)."""
        self.updatePointForces(self.time)
        if self._beganStepAt != self.timeIndex:
            self._check(self.lib.bf_begin_step(self._ctx, float(self.case.deltaT)))
            self._beganStepAt = self.timeIndex
        self._check(self.lib.bf_evolve_begin(self._ctx, C.byref(self.tol)))

    def evolveEnd(self) -> float:
        """Second half (bf_evolve_end): waits, checks convergence as evolve() does, returns the initial residual norm."""
        out = L.bf_step_out()
        rc = self.lib.bf_evolve_end(self._ctx, C.byref(out))
        self.lastStep = out
        self._iOuterCorr = out.nIter
        self.totalIter_ += out.nIter
        self._check(rc)
        return out.initialResidualNorm

    def newton_iteration(self) -> np.ndarray:
        s = (C.c_double * 3)()
        self._check(self.lib.bf_newton_iteration(self._ctx, s))
        return np.array(s[:])

    def loop(self) -> bool:
        """``runTime.loop()``: advance time first, False past endTime."""
        if self.time > self.case.endTime - 0.5 * self.case.deltaT:
            return False
        self.timeIndex += 1
        self.time = self.timeIndex * self.case.deltaT
        return True

    def run(self, nSteps: Optional[int] = None, callback=None):
        """beamFoam.C:73-89 time loop.  Returns per-step iteration counts."""
        iters = []
        while (nSteps is None or len(iters) < nSteps) and self.loop():
            self.evolve()
            iters.append(self._iOuterCorr)
            if callback is not None:
                callback(self)
        return iters
