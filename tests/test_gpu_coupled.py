"""Coupled bundles on the GPU: block-sparse BiCGStab preconditioned by the per-beam block-Thomas
(beam_bicgstab.cuh) against the CPU oracle's restatement of the same algorithm (tests/test_oracle_coupled.py pins that one
to dense LU and to scipy).  The coupling is synthetic and parity with the reference is unpinned (the reference compiles
neither multibeamFvBlockMatrix nor myBlockBiCGStabSolver and ships no contact model)."""
import dataclasses

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam
from test_oracle_coupled import coupled_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme", ["Newmark", "steadyState", "Euler"])
def test_coupled_bundle_matches_oracle(gpu_lib, oracle_lib, scheme):
    case = coupled_case(nBeams=7, n=23, k=5e4, scheme=scheme)
    if scheme == "steadyState":  # the tutorial's 2000 N tip force needs load steps when there is no inertia
        from beamfoam_b200.fvPatchFields import forceBeamDisplacementNR, interpolationTable
        case = dataclasses.replace(case, W_right=forceBeamDisplacementNR(forceSeries=interpolationTable(
            [(0, (0, 0, 0)), (0.02, (200.0, 200.0, 0)), (1, (200.0, 200.0, 0))])))
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    for step in range(12):
        assert g.loop() and o.loop()
        g.evolve()
        o.evolve()
        assert g.iOuterCorr() == o.iOuterCorr(), f"step {step}"
        assert g.lib.bf_linear_iterations(g._ctx) <= 40
        for a, b in zip(g.tip(1), o.tip(1)):
            assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b)), f"step {step}"
    for f in ("W", "Theta", "Q", "M"):
        a, b = g.get_field(f)[0], o.get_field(f)[0]
        assert np.max(np.abs(a - b)) <= 1e-8 * np.max(np.abs(b)) + 1e-12, f
    # the coupling does something: the same bundle without it ends somewhere else
    u = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(case, couplings=None), library=gpu_lib)
    u.run(nSteps=12)
    assert np.max(np.abs(u.tip(1)[0] - g.tip(1)[0])) > 1e-4 * np.max(np.abs(g.tip(1)[0]))
    for m in (g, o, u):
        m.close()


def test_couplings_can_be_removed(gpu_lib, oracle_lib):
    """bf_set_couplings(0) restores the uncoupled kernels: identical to a context that never had any."""
    case = cases.multiBeamDensity(5, 16, jitter=True)
    a = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(case, couplings=[(15, 31, 1e3)]), library=gpu_lib)
    a._check(a.lib.bf_set_couplings(a._ctx, 0, None, None, None))
    b = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    assert a.run(nSteps=5) == b.run(nSteps=5)
    np.testing.assert_array_equal(a.tip(1)[0], b.tip(1)[0])
    a.close()
    b.close()
