"""the run-time selectable beamMomentumContribution sources (Evolve.C:539-571):
morisonDragContribution.C:103-195 and groundContactContribution.C:111-236, with the Hermite-spline mid-point
derivative they evaluate (HermiteSpline.C:83-108, 300-385, 468-513; CTL.C:1263-1427, 1518-1645).

The reference ships no case that uses them, so the oracle is anchored analytically (free-free beam in uniform
motion: the source is a known body force, the response a known acceleration), and the CUDA path is compared with
the oracle.  Two reference properties are replicated on purpose and pinned here:
  * groundContact's normal force is a force per CELL (no factor L), and it enters `source +=` like the drag does,
    i.e. with the sign of a force that pulls the beam further below groundZ (Evolve.C:553-556 adds, the inertia term
    is added with the same sign: m a = loads - contribution);
  * the spline of the reference is built for beam 0 only; here every beam has its own.
"""
import dataclasses

import numpy as np
import pytest

from beamfoam_b200 import BeamGpuError, cases, coupledTotalLagNewtonRaphsonBeam
from beamfoam_b200 import _lib as L
from beamfoam_b200.beamMomentumContribution import groundContactContribution, morisonDragContribution
from beamfoam_b200.fvPatchFields import forceBeamDisplacementNR, interpolationTable, momentBeamRotationNR

Z = (0.0, 0.0, 0.0)


def free_beam(n=12, rotation=None, **kw):
    """A free-free straight beam (zero end forces and moments), Newmark, no gravity."""
    base = cases.freeFlexibleBeam()
    zero = interpolationTable([(0, Z), (1000, Z)])
    return dataclasses.replace(
        base, nSegments=n, W_left=forceBeamDisplacementNR(forceSeries=zero), W_right=forceBeamDisplacementNR(forceSeries=zero),
        Theta_left=momentBeamRotationNR(momentSeries=zero), Theta_right=momentBeamRotationNR(momentSeries=zero),
        initialRotation=rotation, scalingInertiaTensor=1.0, deltaT=1e-3, R=0.44721360, **kw)


def uniform_velocity(m, v):
    m.set_field("U", np.tile(np.asarray(v, dtype=np.float64), (m.nCells, 1)))


def _drag_fixed_point(k, u0, dt):
    """a = -k |u| u with u = u0 + dt a / 2."""
    a = 0.0
    for _ in range(200):
        u = u0 + 0.5 * dt * a
        a = -k * abs(u) * u
    return a


def accel_after_one_step(m):
    m.run(nSteps=1)
    return m.get_field("Accl")[0]


def test_morison_drag_on_a_beam_in_uniform_motion(oracle_lib):
    """U = (ut, un, 0) on a beam along x: dRdScell = e_x, so the source is rhoF R L (Cdt ut^2 e_x + Cdn un^2 e_y) per
    cell and every cell decelerates by that force over rho A L (trapezoidal Newmark: the mean of the step)."""
    Cdn, Cdt, rhoF = 1.2, 0.3, 1000.0
    c = free_beam(momentumContributions=[morisonDragContribution(Cdn, Cdt, rhoF)])
    m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    ut, un = 0.7, -0.4
    uniform_velocity(m, (ut, un, 0.0))
    a = accel_after_one_step(m)
    rhoA = c.rho * c.A
    # at convergence the explicit source is the drag at the NEW velocity u0 + dt (a_old + a)/2, a_old = 0 (Newmark 1/4, 1/2)
    expect = np.array([_drag_fixed_point(rhoF * Cdt * c.R / rhoA, ut, c.deltaT), _drag_fixed_point(rhoF * Cdn * c.R / rhoA, un, c.deltaT), 0.0])
    assert np.allclose(a, expect, rtol=1e-6, atol=1e-9), (a[0], expect)
    assert np.max(np.abs(a - a[0])) < 1e-9  # rigid translation: no cell differs
    m.close()


def test_morison_drag_follows_the_rotated_beam_axis(oracle_lib):
    """Initial configuration rotated 90 degrees about (0 -1 0): the beam axis is e_z, so a velocity along z is tangential."""
    from beamfoam_b200 import axisAngleRotation
    Cdn, Cdt, rhoF = 1.2, 0.3, 1000.0
    c = free_beam(rotation=axisAngleRotation((0, -1, 0), 90.0), momentumContributions=[morisonDragContribution(Cdn, Cdt, rhoF)])
    m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    uniform_velocity(m, (0.0, 0.0, 0.5))
    a = accel_after_one_step(m)
    assert np.allclose(a[:, 2], _drag_fixed_point(rhoF * Cdt * c.R / (c.rho * c.A), 0.5, c.deltaT), rtol=1e-6)
    assert np.max(np.abs(a[:, :2])) < 1e-9
    m.close()


def test_morison_drag_is_undefined_for_steady_state(oracle_lib):
    c = dataclasses.replace(cases.cantilever(nSegments=8), momentumContributions=[morisonDragContribution(1.0, 0.1, 1000.0)], R=0.2)
    with pytest.raises(BeamGpuError) as e:
        coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    assert e.value.status == L.BF_ERR_UNSUPPORTED


def test_ground_contact_normal_force_as_the_reference_applies_it(oracle_lib):
    """A beam at rest, translated to z = -0.05, groundZ = 0: f = 2 R kN (groundZ - z) per CELL; the reference adds it to
    the source, so every cell accelerates with -f/(rho A L) (downwards)."""
    kN = 3.0
    c = free_beam(momentumContributions=[groundContactContribution(kN, 0.0, 0.0, 0.0, 0.0)], initialTranslation=np.array([0.0, 0.0, -0.05]))
    m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    a = accel_after_one_step(m)
    Lc = c.length / c.nSegments
    f = 2 * c.R * kN * 0.05
    assert np.allclose(a[:, 2], -f / (c.rho * c.A * Lc), rtol=1e-3), (a[0], f)
    # above the ground nothing happens
    c2 = dataclasses.replace(c, initialTranslation=np.array([0.0, 0.0, 0.05]))
    m2 = coupledTotalLagNewtonRaphsonBeam(c2, library=oracle_lib)
    assert np.max(np.abs(accel_after_one_step(m2))) < 1e-12
    m.close(); m2.close()


def test_ground_contact_friction_switches_between_stick_and_slip(oracle_lib):
    """Tangential part: -2 R kT Ut while kT 2R |Ut| < mu |fn|, else -mu |fn| UtHat (again added to the source)."""
    kN, kT, mu, depth, ut = 3.0, 0.2, 0.5, 0.05, 0.2
    for kT_, slip in ((kT, False), (50.0, True)):
        c = free_beam(momentumContributions=[groundContactContribution(kN, 0.0, kT_, mu, 0.0)], initialTranslation=np.array([0.0, 0.0, -depth]))
        m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
        uniform_velocity(m, (ut, 0.0, 0.0))
        a = accel_after_one_step(m)
        fn = 2 * c.R * kN * depth
        assert (kT_ * 2 * c.R * ut >= mu * fn) == slip
        ft = -mu * fn if slip else -2 * c.R * kT_ * ut
        Lc = c.length / c.nSegments
        assert np.allclose(a[:, 0], -ft / (c.rho * c.A * Lc), rtol=2e-2), (a[0], ft)
        m.close()


def test_spline_midpoint_derivative_on_a_rolled_up_beam(oracle_lib):
    """tutorials/cantilever at t = 0.25: the prescribed end rotation rolls the beam into a quarter circle in the x-z plane,
    so the unit tangent at arc length s is known; the Hermite-spline mid-point derivative (CTL.C:1263-1427, 1518-1645,
    HermiteSpline.C:83-108) must reproduce it to the O(h^2) of the discretisation."""
    import ctypes as C
    n = 40
    c = dataclasses.replace(cases.cantilever(nSegments=n, deltaT=0.0125), R=0.2,
                            momentumContributions=[groundContactContribution(0.0, 0.0, 0.0, 0.0, -1e30)])  # inert: z never below
    m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    m.run(nSteps=20)
    out = np.zeros((n, 3))
    oracle_lib.oracle_midpoint_derivatives.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    oracle_lib.oracle_midpoint_derivatives.restype = None
    oracle_lib.oracle_midpoint_derivatives(m._ctx, 0, out.ctypes.data)
    W = m.get_field("W")[0]
    pos = W + np.stack([(np.arange(n) + 0.5) * c.length / n, np.zeros(n), np.zeros(n)], axis=1)
    # tangent from the cell-centre positions themselves (central differences): independent of any sign convention
    fd = np.gradient(pos, axis=0)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-12)
    assert np.max(np.abs(out[1:-1] - fd[1:-1])) < 2e-3
    # and it is a quarter circle: the tangent turns by pi/2 over the length
    turn = np.arccos(np.clip(out[0] @ out[-1], -1, 1))
    assert abs(turn - (np.pi / 2) * (n - 1) / n) < 2e-2
    m.close()


def _bundle(nBeams=3, n=16, steps=25):
    """Transient cantilever bundle (Newmark) with both contributions, per-beam translation, ragged lengths."""
    c = cases.multiBeamDensity(nBeams, n, jitter=True)
    c = dataclasses.replace(
        c, nSegments=np.array([n - 3 * (b % 2) for b in range(nBeams)], np.int32), R=0.1, deltaT=2e-3,
        initialTranslation=np.array([[0.0, 0.4 * b, 0.1] for b in range(nBeams)]),
        momentumContributions=[morisonDragContribution(1.2, 0.05, 1000.0), groundContactContribution(2e3, 50.0, 4e2, 0.3, 1.2)])
    return c, steps


def test_contributions_change_the_response_and_keep_newton_converging(oracle_lib):
    c, steps = _bundle()
    a = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    b = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(c, momentumContributions=None), library=oracle_lib)
    ia, ib = a.run(nSteps=steps), b.run(nSteps=steps)
    assert max(ia) < 30 and max(ib) <= 5
    ea, eb = a.beamEnergy(), b.beamEnergy()
    assert np.all(ea[:, 1] < eb[:, 1])  # the drag takes kinetic energy out
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["Newmark", "Euler"])
def test_gpu_contributions_match_oracle(gpu_lib, oracle_lib, scheme):
    c, steps = _bundle(nBeams=67)
    c = dataclasses.replace(c, d2dt2Scheme=scheme)
    g = coupledTotalLagNewtonRaphsonBeam(c, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    gi, oi = g.run(nSteps=steps), o.run(nSteps=steps)
    assert gi == oi
    for x, y in zip(g.tip(1), o.tip(1)):
        assert np.max(np.abs(x - y)) <= 1e-9 * np.max(np.abs(y))
    for name in ("U", "Accl"):
        x, y = g.get_field(name)[0], o.get_field(name)[0]
        assert np.max(np.abs(x - y)) <= 1e-8 * np.max(np.abs(y)), name
    g.close(); o.close()


@pytest.mark.gpu
def test_gpu_contributions_with_point_forces_and_steady_ground(gpu_lib, oracle_lib):
    """groundContact is defined for steadyState too (U = 0), and shares the per-cell source array with the point forces."""
    base = dataclasses.replace(cases.largeDeflectionCantilever(tipForce=0.2, nSegments=24), R=0.0002,
                               initialTranslation=np.array([0.0, 0.0, 0.0]),
                               momentumContributions=[groundContactContribution(5.0, 0.0, 0.0, 0.0, -0.01)],
                               pointForces=[(11, 0.3, (0.0, 0.0, -0.05))])
    g = coupledTotalLagNewtonRaphsonBeam(base, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(base, library=oracle_lib)
    assert g.run() == o.run()
    for x, y in zip(g.tip(1), o.tip(1)):
        assert np.max(np.abs(x - y)) <= 1e-7 * np.max(np.abs(y))  # tolerance of this case: tests/test_gpu_parity.py
    g.close(); o.close()
