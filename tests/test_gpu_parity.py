"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same cases.

Tolerances (north-star): tip displacement and rotation within 1e-10 RELATIVE, identical
Newton iteration counts per step.  "Relative" is taken against the largest magnitude of
the compared quantity over the beam's tips at that step (max-norm), the way a tip-displacement
history is judged.
"""
import dataclasses

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam

pytestmark = pytest.mark.gpu

TIP_RTOL = 1e-10
ALL_FIELDS = ("W", "Theta", "Lambda", "DW", "DTheta", "Lambdaf", "Gamma", "K", "Q", "M")
NEWMARK_FIELDS = ("U", "Accl", "Omega", "dotOmega")


def rel(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def run_pair(case, gpu_lib, oracle_lib, nSteps, tip_rtol=TIP_RTOL, field_rtol=1e-8, check_fields=True):
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    worst = 0.0
    for step in range(nSteps):
        assert g.loop() and o.loop()
        g.evolve()
        o.evolve()
        assert g.iOuterCorr() == o.iOuterCorr(), f"step {step}: Newton iterations {g.iOuterCorr()} vs {o.iOuterCorr()}"
        tips = [(g.tip(end), o.tip(end)) for end in (0, 1)]
        for k, nm in ((0, "W"), (1, "Theta")):
            # max-norm over the tip values of this step: an end that has barely moved yet (the far end of
            # the free-free beam holds 1e-11 while the loaded end holds 1e-2) is judged on the scale of
            # the beam's tip motion, not on its own rounding-level magnitude
            scale = max(float(np.max(np.abs(tips[end][1][k]))) for end in (0, 1))
            for end in (0, 1):
                a, b = tips[end][0][k], tips[end][1][k]
                if np.max(np.abs(b)) < 1e-13:  # a clamped end: both must be (numerically) zero
                    assert np.max(np.abs(a)) < 1e-13
                    continue
                r = float(np.max(np.abs(a - b)) / max(scale, 1e-300))
                worst = max(worst, r)
                assert r <= tip_rtol, f"step {step} end {end} {nm}: rel diff {r:.3e}"
        # the step's initial residual carries the (1e-10-level) state difference times the
        # stiffness, so it is only checked loosely; the state itself is checked above and below
        assert abs(g.lastStep.initialResidualNorm - o.lastStep.initialResidualNorm) <= \
            1e-6 * max(o.lastStep.initialResidualNorm, 1e-300)
    if check_fields:
        # Fields that are physically ~0 (e.g. Gamma in pure bending) hold only rounding noise, so
        # every field is judged against max|field| PLUS a floor of 1e-13 x its natural scale.
        c = case
        Lb = float(np.max(c.per_beam(c.length)))
        dZ = float(np.min(c.per_beam(c.length) / c.per_beam(c.nSegments)))
        EA = float(np.max(c.per_beam(c.E) * c.per_beam(c.A)))
        EI = float(np.max(c.per_beam(c.E) * np.maximum(c.per_beam(c.Iyy), c.per_beam(c.Izz))))
        dt = c.deltaT
        natural = {"W": Lb, "DW": Lb, "Theta": 1.0, "DTheta": 1.0, "Lambda": 1.0, "Lambdaf": 1.0, "Gamma": 1.0,
                   "K": 1.0 / dZ, "Q": EA, "M": EI / dZ, "U": Lb / dt, "Accl": Lb / dt ** 2, "Omega": 1.0 / dt,
                   "dotOmega": 1.0 / dt ** 2}
        names = ALL_FIELDS + (NEWMARK_FIELDS if case.d2dt2Scheme == "Newmark" else ())
        for name in names:
            ref = o.get_field(name)
            # the last increments DW/DTheta are converged-to-noise: judge them on the scale of W/Theta
            mag = o.get_field({"DW": "W", "DTheta": "Theta"}.get(name, name))[0]
            scale = field_rtol * float(np.max(np.abs(mag))) + 1e-13 * natural[name]
            for a, b, part in zip(g.get_field(name), ref, ("internal", "left", "right")):
                if a is None or not a.size:
                    continue
                d = float(np.max(np.abs(a - b)))
                assert d <= scale, f"field {name} ({part}): abs diff {d:.3e} > {scale:.3e}"
    g.close()
    o.close()
    return worst


def test_cantilever_every_cell(gpu_lib, oracle_lib):
    """configs[0]: tutorials/cantilever, 1 x 50 cells, all 100 load steps, every cell."""
    run_pair(cases.cantilever(), gpu_lib, oracle_lib, nSteps=100)


def test_follower_force(gpu_lib, oracle_lib):
    run_pair(cases.cantileverFollowerForce(), gpu_lib, oracle_lib, nSteps=200)


def test_follower_force_long_beam(gpu_lib, oracle_lib):
    """configs[1] physics at reduced length for the oracle's sake: 2000 segments, 5 steps."""
    run_pair(cases.cantileverFollowerForce(nSegments=2000), gpu_lib, oracle_lib, nSteps=5)


@pytest.mark.parametrize("F", [0.0, 0.294, 0.588])
def test_large_deflection(gpu_lib, oracle_lib, F):
    # very slender strip (EA/EI ~ 7.5e7, Iyy/Izz ~ 4e3): the rotational 3x3 pivot block is ill-conditioned and the 6x6 blocks
    # go through the pivoted LU (smallDet); measured difference to the oracle 1e-15 (round 1, without the pivoting: 1e-8)
    run_pair(cases.largeDeflectionCantilever(F), gpu_lib, oracle_lib, nSteps=1)


def test_dynamic_cantilever(gpu_lib, oracle_lib):
    run_pair(cases.dynamicCantilever(), gpu_lib, oracle_lib, nSteps=200)


def test_multibeam_density(gpu_lib, oracle_lib):
    """tutorials/3DdynamicCantileverMultiBeamDensity verbatim (3 beams x 40 cells)."""
    run_pair(cases.multiBeamDensity(3, 40), gpu_lib, oracle_lib, nSteps=100)


def test_free_flexible_beam(gpu_lib, oracle_lib):
    run_pair(cases.freeFlexibleBeam(), gpu_lib, oracle_lib, nSteps=300)


def test_spring_in_series(gpu_lib, oracle_lib):
    run_pair(cases.beamAndSpringInSeries(), gpu_lib, oracle_lib, nSteps=10)


def test_cantilever_spring(gpu_lib, oracle_lib):
    run_pair(cases.cantileverSpring(), gpu_lib, oracle_lib, nSteps=1)


def test_ragged_bundle_with_jitter(gpu_lib, oracle_lib):
    """Ragged nSegments (padding path) + per-beam load jitter on a 70-beam bundle."""
    case = cases.multiBeamDensity(70, 24, jitter=True)
    nseg = 8 + (np.arange(70) * 7) % 23
    case = dataclasses.replace(case, nSegments=nseg)
    run_pair(case, gpu_lib, oracle_lib, nSteps=20)


def test_bundle_4096_subset_matches_oracle(gpu_lib, oracle_lib):
    """configs[2] shape (4096 beams x 200 cells on the GPU): a 64-beam window of the oracle
    must match the same beams inside the big GPU bundle (beams are independent; only the
    global convergence norm couples them, so compare after a FIXED number of iterations)."""
    big = cases.multiBeamDensity(4096, 200, jitter=True)
    g = coupledTotalLagNewtonRaphsonBeam(big, library=gpu_lib, storeIncrements=False)
    o = coupledTotalLagNewtonRaphsonBeam(big, library=oracle_lib, beamRange=(1000, 1064))
    for m in (g, o):
        assert m.loop()
        m.updateCoeffs(m.time)
        m.lib.bf_begin_step(m._ctx, big.deltaT)
        for _ in range(4):
            m.newton_iteration()
    gw, gt = g.tip(1)
    ow, ot = o.tip(1)
    assert rel(gw[1000:1064], ow) <= TIP_RTOL
    assert rel(gt[1000:1064], ot) <= TIP_RTOL
    g.close()
    o.close()


# ---------------------------------------------------------------------------------------------
# SURVEY 8a row a4, the two parts no tutorial exercises: constant/pointForces and the Euler d2dt2Scheme
# ---------------------------------------------------------------------------------------------
def test_point_forces_steady_and_newmark(gpu_lib, oracle_lib):
    from beamfoam_b200.fvPatchFields import interpolationTable
    # ragged 5-beam bundle, point forces on first / last / interior cells, both sides of a cell, a time series
    base = cases.multiBeamDensity(5, 12)
    nseg = np.array([12, 7, 9, 12, 5])
    start = np.concatenate([[0], np.cumsum(nseg)])
    ser = interpolationTable([(0, (0, 0, 0)), (0.01, (50.0, -20.0, 10.0)), (1, (50.0, -20.0, 10.0))])
    pfs = [(int(start[0]), -1.0, (30.0, 0, 5.0)), (int(start[0]) + 5, 0.3, ser), (int(start[0]) + 5, -0.7, (0, 12.0, 0)),
           (int(start[1]) + 6, 1.0, (0, -40.0, 0)), (int(start[2]) + 4, 0.0, (10.0, 10.0, 10.0)),
           (int(start[4]), 0.5, ser), (int(start[4]) + 4, -0.25, (0, 0, -25.0))]
    for scheme in ("Newmark", "steadyState"):
        case = dataclasses.replace(base, nSegments=nseg, pointForces=pfs, d2dt2Scheme=scheme)
        if scheme == "steadyState":  # the tutorial's 2000 N tip force needs load steps when there is no inertia
            case = dataclasses.replace(case, W_right=type(case.W_right)(forceSeries=interpolationTable(
                [(0, (0, 0, 0)), (0.02, (200.0, 200.0, 0)), (1, (200.0, 200.0, 0))])))
        run_pair(case, gpu_lib, oracle_lib, nSteps=20, tip_rtol=1e-10)


def test_euler_scheme(gpu_lib, oracle_lib):
    case = dataclasses.replace(cases.multiBeamDensity(7, 16, jitter=True), d2dt2Scheme="Euler")
    worst = run_pair(case, gpu_lib, oracle_lib, nSteps=60)
    assert worst <= 1e-10
    # a change of time step between steps (old-time levels are per step, not per dt)
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    for m in (g, o):
        m.run(nSteps=3)
        m.case = dataclasses.replace(m.case, deltaT=2.5e-3)
        m.run(nSteps=3)
    for a, b in zip(g.tip(1), o.tip(1)):
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    for f in ("U", "Accl", "Omega"):
        a, b = g.get_field(f)[0], o.get_field(f)[0]
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), 1e-3), f


@pytest.mark.parametrize("name", ["cantileverFollowerForce", "multiBeamDensity"])
def test_evolve_in_two_halves_is_evolve_on_the_gpu(gpu_lib, name):
    """bf_evolve_begin / bf_evolve_end with the next level's BC values prepared in between: same iteration counts and
    bit-identical tips as bf_evolve; calls that would queue work behind the open loop are refused."""
    import ctypes as C
    from beamfoam_b200 import _lib as L
    from test_abi_and_host import run_two_halves
    case = cases.cantileverFollowerForce() if name == "cantileverFollowerForce" else cases.multiBeamDensity(70, 37, jitter=True)
    a = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    b = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    assert a.run(nSteps=15) == run_two_halves(b, 15)
    for x, y in zip(a.tip(1), b.tip(1)):
        assert np.array_equal(x, y)
    assert gpu_lib.bf_evolve_begin(b._ctx, C.byref(b.tol)) == L.BF_OK
    s = (C.c_double * 3)()
    assert gpu_lib.bf_newton_iteration(b._ctx, s) == L.BF_ERR_INVALID
    assert gpu_lib.bf_begin_step(b._ctx, 1.0) == L.BF_ERR_INVALID
    buf = np.zeros((b._B, 3))
    assert gpu_lib.bf_get_patch_pair(b._ctx, 1, buf.ctypes.data, buf.ctypes.data) == L.BF_ERR_INVALID
    assert gpu_lib.bf_evolve_begin(b._ctx, C.byref(b.tol)) == L.BF_ERR_INVALID
    out = L.bf_step_out()
    gpu_lib.bf_evolve_end(b._ctx, C.byref(out))
    assert gpu_lib.bf_evolve_end(b._ctx, C.byref(out)) == L.BF_ERR_INVALID
    a.close()
    b.close()
