"""GPU parity of the cell-parallel path for few long beams (beam_long.cuh: assembled block rows + block parallel
cyclic reduction, BASELINE configs[1]) against the CPU oracle.  BEAMGPU_LONG_BEAMS=1 forces the path on bundles
that would otherwise take the fused sweeps, so every boundary-condition class and time scheme goes through it."""
import dataclasses

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam
from beamfoam_b200.fvPatchFields import (fixedDisplacement, fixedRotation, fixedValue, followerForceBeamDisplacementNR,
                                         forceBeamDisplacementNR, interpolationTable,
                                         linearSpringForceBeamDisplacementNR, momentBeamRotationNR)
from test_gpu_parity import run_pair

pytestmark = pytest.mark.gpu


@pytest.fixture
def long_path(monkeypatch):
    monkeypatch.setenv("BEAMGPU_LONG_BEAMS", "1")


def test_config1_follower_beam_10000_cells(gpu_lib, oracle_lib):
    """BASELINE configs[1]: cantileverFollowerForce with nSegments 10000 (the path is chosen automatically).  20 load
    steps: without the refinement pass of the cyclic reduction the Newton loop needs 5-6 iterations from step 3 on
    where the reference-shaped pivoted LU of the oracle needs 4."""
    # tips at 1e-10 (measured 3e-15); the field check is looser only because Gamma divides W differences of 1e-15 by dZ = 0.01
    run_pair(cases.cantileverFollowerForce(nSegments=10000), gpu_lib, oracle_lib, nSteps=20, field_rtol=1e-7)


def test_cantilever_roll_up(gpu_lib, oracle_lib, long_path):
    run_pair(cases.cantilever(), gpu_lib, oracle_lib, nSteps=100)


def test_newmark_bundle_ragged(gpu_lib, oracle_lib, long_path):
    case = cases.multiBeamDensity(9, 24, jitter=True)
    case = dataclasses.replace(case, nSegments=8 + (np.arange(9) * 7) % 17)
    run_pair(case, gpu_lib, oracle_lib, nSteps=40)


def test_free_flexible_beam(gpu_lib, oracle_lib, long_path):
    run_pair(cases.freeFlexibleBeam(), gpu_lib, oracle_lib, nSteps=100)


def test_euler_and_point_forces(gpu_lib, oracle_lib, long_path):
    base = dataclasses.replace(cases.multiBeamDensity(3, 10), d2dt2Scheme="Euler",
                               pointForces=[(4, 0.5, (20.0, 0, -10.0)), (10, -1.0, (0, 15.0, 0)), (29, 1.0, (5.0, 5.0, 5.0))])
    run_pair(base, gpu_lib, oracle_lib, nSteps=30)


def test_every_bc_class(gpu_lib, oracle_lib, long_path):
    ramp = lambda v: interpolationTable([(0, (0, 0, 0)), (1, v)])  # noqa: E731
    W_right = [forceBeamDisplacementNR(forceSeries=ramp((0, 30, -40))),
               followerForceBeamDisplacementNR(followerForceSeries=ramp((0, 0, 8))),
               linearSpringForceBeamDisplacementNR(springStiffness=0.05, springDirection=(0, 1, 1),
                                                   offsetForceSeries=ramp((0, 0.5, 0))),
               fixedDisplacement(displacementSeries=ramp((-0.2, 0.3, 0))),
               forceBeamDisplacementNR(force=(0, 0, 0)),
               forceBeamDisplacementNR(forceSeries=ramp((5, 0, 20)))]
    Th_right = [momentBeamRotationNR(momentSeries=ramp((0, 10, 0))), momentBeamRotationNR(moment=(0, 0, 0)),
                momentBeamRotationNR(moment=(0, 0, 0)), fixedRotation(thetaSeries=ramp((0.1, -0.3, 0.2))),
                fixedRotation(thetaSeries=ramp((0, -1.0, 0))), fixedValue(variable="Theta")]
    W_left = [fixedValue(variable="W")] * 2 + [fixedDisplacement(displacementSeries=ramp((0, 0, 0.05))),
              fixedValue(variable="W")] + [fixedDisplacement(displacementSeries=ramp((0, 0, 0.1))),
                                               fixedValue(variable="W")]
    Th_left = [fixedValue(variable="Theta")] * 5 + [fixedRotation(thetaSeries=ramp((0, 0, 0.2)))]
    case = dataclasses.replace(cases.cantilever(deltaT=0.1), nBeams=6, nSegments=[12, 17, 9, 14, 20, 1], W_left=W_left,
                               W_right=W_right, Theta_left=Th_left, Theta_right=Th_right, nCorrectors=400)
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    assert g.run(nSteps=5) == o.run(nSteps=5)
    for end in (0, 1):
        for a, b in zip(g.tip(end), o.tip(end)):
            assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(b)), 1e-3)
    for f in ("Q", "M", "Lambdaf", "K"):
        for a, b in zip(g.get_field(f), o.get_field(f)):
            if a.size:
                assert np.max(np.abs(a - b)) <= 1e-8 * max(np.max(np.abs(b)), 1.0), f


def test_momentum_contributions(gpu_lib, oracle_lib, long_path):
    """The per-cell source array of beam_momentum_contributions feeds long_assemble as it feeds the forward sweep."""
    from beamfoam_b200.beamMomentumContribution import groundContactContribution, morisonDragContribution
    case = dataclasses.replace(cases.multiBeamDensity(4, 18, jitter=True), R=0.1, deltaT=2e-3,
                               initialTranslation=np.array([0.0, 0.0, 0.1]),
                               momentumContributions=[morisonDragContribution(1.2, 0.05, 1000.0),
                                                      groundContactContribution(2e3, 50.0, 4e2, 0.3, 1.2)])
    run_pair(case, gpu_lib, oracle_lib, nSteps=20)
