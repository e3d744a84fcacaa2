"""Coupled bundles on the CPU oracle: the restated BiCGStab of myBlockBiCGStabSolver.C:79-189,
preconditioned by the per-beam banded LU, on the multibeamFvBlockMatrix-shaped system (per-beam block-tridiagonals plus
inter-beam 6x6 blocks).  The reference compiles neither and ships no contact model, so the coupling is a synthetic linear
spring and parity with the reference is UNPINNED; what is pinned here is the linear algebra: the oracle's Krylov solve
against numpy's dense LU and scipy's BiCGStab on the matrix the oracle itself assembled, and the physics of two beams
tied by a spring."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam


def coupled_case(nBeams=4, n=12, k=5e4, scheme="Newmark"):
    """Neighbouring beams of the multi-beam bundle tied at a third and at the tip; loads differ per beam (jitter + rho)."""
    pairs = []
    for b in range(nBeams - 1):
        pairs.append((b * n + n - 1, (b + 1) * n + n - 1, k))
        pairs.append((b * n + n // 3, (b + 1) * n + n // 3 + 1, 0.3 * k))
    return dataclasses.replace(cases.multiBeamDensity(nBeams, n, jitter=True), couplings=pairs, d2dt2Scheme=scheme,
                               name=f"coupled_{nBeams}x{n}")


def test_krylov_solve_equals_dense_lu_and_scipy(oracle_lib):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    case = coupled_case()
    m = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib, storeIncrements=True)
    m.run(nSteps=2)
    assert m.loop()
    m.updateCoeffs(m.time)
    m._check(m.lib.bf_begin_step(m._ctx, float(case.deltaT)))
    m.newton_iteration()   # first iteration of the step done: assemble the SECOND one (no predictor in the way)
    n = 6 * m.nCells
    A, b = np.zeros((n, n)), np.zeros(n)
    m.lib.bf_oracle_dense_system.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    m._check(m.lib.bf_oracle_dense_system(m._ctx, A.ctypes.data, b.ctypes.data))
    # off-tridiagonal blocks exist, exactly where the pairs are
    blocks = np.abs(A.reshape(m.nCells, 6, m.nCells, 6)).max(axis=(1, 3)) > 0
    off = {(i, j) for i, j in zip(*np.nonzero(blocks)) if abs(i - j) > 1}
    want = {(a, b_) for a, b_, _ in case.couplings} | {(b_, a) for a, b_, _ in case.couplings}
    assert off == want
    m.newton_iteration()
    x = np.concatenate([m.get_field("DW")[0], m.get_field("DTheta")[0]], axis=1).ravel()
    its = m.lib.bf_linear_iterations(m._ctx)
    assert 1 <= its <= 25, its   # the block-tridiagonal part dominates: the preconditioned iteration converges in a few steps
    xd = np.linalg.solve(A, b)
    assert np.max(np.abs(x - xd)) <= 1e-9 * np.max(np.abs(xd))
    # scipy's BiCGStab with the same (block-tridiagonal) preconditioner
    tri = A.copy()
    for a, b_, _ in case.couplings:
        tri[6 * a:6 * a + 6, 6 * b_:6 * b_ + 6] = 0
        tri[6 * b_:6 * b_ + 6, 6 * a:6 * a + 6] = 0
    lu = spla.splu(sp.csc_matrix(tri))
    xs, info = spla.bicgstab(sp.csr_matrix(A), b, M=spla.LinearOperator((n, n), lu.solve), rtol=1e-13, atol=0, maxiter=200)
    assert info == 0
    assert np.max(np.abs(x - xs)) <= 1e-8 * np.max(np.abs(xs))
    m.close()


def test_stiff_tie_makes_two_beams_move_together(oracle_lib):
    """Two identical beams, one loaded, tips tied: a soft tie leaves the unloaded beam almost at rest, a stiff tie (k = 1e5, 130 times
    the tip stiffness 3 EI / L^3 = 7.6e2 N/m) shares the load: both tips move by half of the free deflection."""
    base = dataclasses.replace(cases.multiBeamDensity(2, 20), d2dt2Scheme="steadyState", rho=1000.0, deltaT=1.0, endTime=1.0)
    from beamfoam_b200.fvPatchFields import forceBeamDisplacementNR
    load = [forceBeamDisplacementNR(force=(0.0, 20.0, 0.0)), forceBeamDisplacementNR(force=(0.0, 0.0, 0.0))]
    tips = {}
    for k in (0.0, 1e-2, 1e5):
        c = dataclasses.replace(base, W_right=load, couplings=[(19, 39, k)] if k else None)
        m = coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
        m.run()
        tips[k] = m.tip(1)[0][:, 1].copy()
        m.close()
    free = tips[0.0][0]
    assert abs(tips[0.0][1]) < 1e-12 and free > 1e-3
    assert abs(tips[1e-2][1]) < 1e-3 * free
    # the tie acts between the last CELL centres (half a cell from the tips): equal to 1 %
    assert abs(tips[1e5][0] - tips[1e5][1]) < 2e-2 * free
    assert abs(tips[1e5][0] - 0.5 * free) < 2e-2 * free


def test_pairs_are_validated(oracle_lib):
    c = dataclasses.replace(cases.multiBeamDensity(2, 5), couplings=[(0, 10, 1.0)])
    with pytest.raises(Exception):
        coupledTotalLagNewtonRaphsonBeam(c, library=oracle_lib)
