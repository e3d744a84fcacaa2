"""Oracle sanity anchors for the two remaining pieces of SURVEY 8a row a4: constant/pointForces
(Evolve.C:211-271) and the Euler d2dt2Scheme (Evolve.C:385-400, 490-513, 791-795, 864-869).  The reference
ships no case that uses either, so the anchors are analytic / cross-scheme."""
import dataclasses

import numpy as np

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam
from beamfoam_b200.fvPatchFields import forceBeamDisplacementNR, momentBeamRotationNR, interpolationTable


def clean_cantilever(n=40, **kw):
    """Straight cantilever, no self-weight, free right end (zero force, zero moment), one steady step."""
    base = cases.cantilever(nSegments=n)
    return dataclasses.replace(base, W_right=forceBeamDisplacementNR(force=(0, 0, 0)),
                               Theta_right=momentBeamRotationNR(moment=(0, 0, 0)), deltaT=1.0, endTime=1.0, **kw)


def test_point_force_small_load_matches_euler_bernoulli(oracle_lib):
    n = 40
    c = clean_cantilever(n)
    L, EI = c.length, c.E * c.Izz
    F = -1e-3  # small: linear regime
    for cell, zeta in ((n // 2, 0.0), (n // 4, 0.6), (3 * n // 4, -0.4), (n - 1, 1.0), (0, -1.0)):
        m = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(c, pointForces=[(cell, zeta, (0, F, 0))]),
                                             library=oracle_lib)
        m.run()
        a = (cell + 0.5 + 0.5 * zeta) * L / n  # position of the load along the beam
        w_tip = F * a * a * (3 * L - a) / (6 * EI) + F * a / (c.G * c.A)  # bending + shear (CQ = diag(EA, GA, GA))
        got = m.tip(1)[0][0][1]
        if a == 0.0:  # load on the clamp: only the O(h^2) residue of a force + couple on the first cell remains
            assert abs(got) < 1e-2 * abs(F * L ** 3 / (3 * EI))
        else:
            assert abs(got - w_tip) <= 1e-3 * abs(w_tip), (cell, zeta, got, w_tip)
        m.close()


def test_point_force_at_tip_equals_end_force(oracle_lib):
    """zeta = +1 on the last cell puts the force on the right patch face: statically the same load as
    forceBeamDisplacementNR there, also at large deflection."""
    n = 50
    c = dataclasses.replace(clean_cantilever(n), deltaT=0.05, divergenceTol=1e12)
    ramp = interpolationTable([(0, (0, 0, 0)), (1, (0.0, -3.0, 1.2))])  # EI = 100, L = 10: F L^2/EI = 3
    a = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(c, W_right=forceBeamDisplacementNR(forceSeries=ramp)), library=oracle_lib)
    b = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(c, pointForces=[(n - 1, 1.0, ramp)]), library=oracle_lib)
    a.run(); b.run()
    wa, wb = a.tip(1)[0][0], b.tip(1)[0][0]
    assert np.linalg.norm(wa) > 0.5  # genuinely large deflection (L = 10)
    assert np.max(np.abs(wa - wb)) <= 1e-5 * np.max(np.abs(wa))


def test_point_force_series_and_superposition(oracle_lib):
    """Two point forces on one cell (both sides) equal the sum of the separate small-load solutions; the force
    comes from an interpolationTable evaluated at the step's time."""
    n = 30
    c = dataclasses.replace(clean_cantilever(n), deltaT=0.5)
    ser = interpolationTable([(0, (0, 0, 0)), (1, (0, 2e-3, 0))])
    def tip(pfs, steps):
        m = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(c, pointForces=pfs), library=oracle_lib)
        m.run(nSteps=steps)
        return m.tip(1)[0][0].copy()
    both = tip([(10, 0.5, ser), (10, -0.5, (0, 0, -1e-3))], 1)
    one = tip([(10, 0.5, (0, 1e-3, 0))], 1)   # the series at t = 0.5
    two = tip([(10, -0.5, (0, 0, -1e-3))], 1)
    assert np.max(np.abs(both - (one + two))) <= 1e-6 * np.max(np.abs(both))


def test_euler_scheme_relaxes_to_the_static_solution(oracle_lib):
    """Backward Euler is dissipative: with a constant load and a time step far above the first period the
    transient dies out and the state is the steadyState solution of the same load."""
    dyn = cases.dynamicCantilever(nSegments=10, loadScale=0.02)
    static = dataclasses.replace(dyn, d2dt2Scheme="steadyState")
    euler = dataclasses.replace(dyn, d2dt2Scheme="Euler", deltaT=50.0, endTime=1e9)
    s = coupledTotalLagNewtonRaphsonBeam(static, library=oracle_lib)
    e = coupledTotalLagNewtonRaphsonBeam(euler, library=oracle_lib)
    s.run(nSteps=1)
    e.run(nSteps=40)
    ws, we = s.tip(1)[0][0], e.tip(1)[0][0]
    assert np.max(np.abs(ws - we)) <= 1e-6 * np.max(np.abs(ws))
    U = e.get_field("U")[0]
    assert np.max(np.abs(U)) <= 1e-8


def test_euler_scheme_tracks_newmark_at_small_time_steps(oracle_lib):
    """First-order Euler and second-order Newmark integrate the same equations of motion: over a short window
    at small dt the tip histories agree to the first-order error of Euler."""
    nm = cases.dynamicCantilever(nSegments=10, deltaT=2e-4, loadScale=0.2)
    eu = dataclasses.replace(nm, d2dt2Scheme="Euler")
    a = coupledTotalLagNewtonRaphsonBeam(nm, library=oracle_lib)
    b = coupledTotalLagNewtonRaphsonBeam(eu, library=oracle_lib)
    a.run(nSteps=100); b.run(nSteps=100)
    wa, wb = a.tip(1)[0][0], b.tip(1)[0][0]
    assert np.max(np.abs(wa)) > 1e-3
    assert np.max(np.abs(wa - wb)) <= 0.15 * np.max(np.abs(wa))
    # the velocity is the backward difference of W by construction (Evolve.C:793)
    W0 = b.get_field("W")[0].copy()
    b.run(nSteps=1)
    np.testing.assert_allclose(b.get_field("U")[0], (b.get_field("W")[0] - W0) / eu.deltaT, rtol=0, atol=1e-9)


# ---------------------------------------------------------------------------------------------
# bf_beam_energy (functionObjects/beamEnergyData/beamEnergyData.C:84-215)
# ---------------------------------------------------------------------------------------------
def test_beam_energy_static_equals_work_of_the_load(oracle_lib):
    """Small tip load on a cantilever: strain energy = 1/2 F w (Clapeyron), no kinetic energy."""
    F = -1e-3
    c = dataclasses.replace(clean_cantilever(40), W_right=forceBeamDisplacementNR(force=(0, F, 0)))
    m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
    m.run()
    e = m.beamEnergy()
    assert e.shape == (1, 3) and e[0, 1] == 0.0 and e[0, 2] == 0.0
    assert abs(e[0, 0] - 0.5 * F * m.tip(1)[0][0][1]) <= 2e-3 * e[0, 0]


def test_beam_energy_newmark_balance(oracle_lib):
    """Constant dead load from rest: internal + kinetic energy = work of the load, F . w_tip, which the
    average-acceleration Newmark scheme preserves up to the nonlinearity."""
    m = coupledTotalLagNewtonRaphsonBeam(cases.multiBeamDensity(3, 20), library=oracle_lib)
    for _ in range(60):
        m.run(nSteps=1)
    e = m.beamEnergy()
    work = (m.tip(1)[0] * np.array([2000.0, 2000.0, 0.0])).sum(axis=1)
    assert np.all(e > 0)
    np.testing.assert_allclose(e.sum(axis=1), work, rtol=2e-2)
