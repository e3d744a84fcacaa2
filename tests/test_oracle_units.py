"""Unit checks of the oracle's helpers against closed forms, and of its linear solve against
SuperLU with COLAMD ordering (the nearest available stand-in for Eigen::SparseLU<COLAMD>)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial.transform import Rotation

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam

dp = C.POINTER(C.c_double)


def _call(lib, name, a, n_out):
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros(n_out)
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [dp, dp]
    fn(a.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


def test_spin_tensor(oracle_lib):
    rng = np.random.default_rng(1)
    for _ in range(10):
        v, w = rng.normal(size=3), rng.normal(size=3)
        S = _call(oracle_lib, "oracle_spinTensor", v, 9).reshape(3, 3)
        np.testing.assert_allclose(S @ w, np.cross(v, w), atol=1e-15)
        np.testing.assert_allclose(S, -S.T)


def test_rotation_matrix_rodrigues(oracle_lib):
    rng = np.random.default_rng(2)
    for scale in (1e-9, 1e-3, 0.5, 3.0):
        th = rng.normal(size=3) * scale
        R = _call(oracle_lib, "oracle_rotationMatrix", th, 9).reshape(3, 3)
        np.testing.assert_allclose(R, Rotation.from_rotvec(th).as_matrix(), atol=2e-9 if scale < 1e-6 else 1e-13)
    R0 = _call(oracle_lib, "oracle_rotationMatrix", np.zeros(3), 9).reshape(3, 3)
    np.testing.assert_allclose(R0, np.eye(3), atol=0)  # phi = SMALL: no 0/0


def test_tensor_inverse(oracle_lib):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(3, 3)) + 3 * np.eye(3)
    Ai = _call(oracle_lib, "oracle_inv", A.ravel(), 9).reshape(3, 3)
    np.testing.assert_allclose(Ai @ A, np.eye(3), atol=1e-13)


@pytest.mark.parametrize("mk,tol", [(lambda: cases.cantileverFollowerForce(nSegments=12), 1e-9),
                                    (lambda: cases.multiBeamDensity(3, 9), 1e-9),
                                    # slender strip: cond(A) ~ 1e9, two backward-stable solves agree to cond*eps
                                    (lambda: cases.largeDeflectionCantilever(0.3, nSegments=16), 1e-6)])
def test_linear_solve_vs_superlu_colamd(oracle_lib, mk, tol):
    """d/l/u/source of an iteration, assembled into the reference's global matrix layout
    (BlockEigenSolverOF.C:60-175), solved by SuperLU(COLAMD): same x as the oracle's banded LU."""
    m = coupledTotalLagNewtonRaphsonBeam(mk(), library=oracle_lib)
    m.loop()
    m.updateCoeffs(m.time)
    m.lib.bf_begin_step(m._ctx, m.case.deltaT)
    for _ in range(2):
        m.newton_iteration()
    N, F = m.nCells, m.nInternalFaces
    d, l, u = np.zeros((N, 6, 6)), np.zeros((F, 6, 6)), np.zeros((F, 6, 6))
    src, x = np.zeros((N, 6)), np.zeros((N, 6))
    fn = m.lib.oracle_get_system
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    fn(m._ctx, d.ctypes.data, l.ctypes.data, u.ctypes.data, src.ctypes.data, x.ctypes.data)
    own, nei = [], []
    for b, n in enumerate(m.nSeg):
        for i in range(1, n):
            own.append(m.cellStart[b] + i - 1)
            nei.append(m.cellStart[b] + i)
    A = sp.lil_matrix((6 * N, 6 * N))
    for c in range(N):
        A[6 * c:6 * c + 6, 6 * c:6 * c + 6] = d[c]
    for f, (o, n_) in enumerate(zip(own, nei)):
        A[6 * o:6 * o + 6, 6 * n_:6 * n_ + 6] = u[f]
        A[6 * n_:6 * n_ + 6, 6 * o:6 * o + 6] = l[f]
    xs = spla.splu(sp.csc_matrix(A), permc_spec="COLAMD").solve(src.ravel())
    assert np.max(np.abs(xs - x.ravel())) <= tol * np.max(np.abs(xs))
    # backward stability of the oracle's banded LU: ||A x - b|| <= 1e-13 (||A|| ||x|| + ||b||)
    Ac = sp.csc_matrix(A)
    r = Ac @ x.ravel() - src.ravel()
    assert np.linalg.norm(r, np.inf) <= 1e-13 * (spla.norm(Ac, np.inf) * np.linalg.norm(x.ravel(), np.inf)
                                                 + np.linalg.norm(src.ravel(), np.inf))
    m.close()
