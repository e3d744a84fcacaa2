"""CPU checks of the boundary: the CUDA library loads and exports every symbol of
include/beamgpu.h (no compute calls without a GPU), fails loudly without a device, and the
host-side logic (time series, sharding, argument checking) behaves like the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from beamfoam_b200 import BeamGpuError, _lib as L, cases, coupledTotalLagNewtonRaphsonBeam, shard_range
from beamfoam_b200.fvPatchFields import interpolationTable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "beamgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", txt)) - {"bf_reduce_fn"})


def test_header_and_binding_agree():
    assert header_symbols() == sorted(L.EXPORTS)


def test_cuda_library_exports_every_symbol(gpu_lib):
    for name in header_symbols():
        assert hasattr(gpu_lib, name), name
    assert gpu_lib.bf_backend() == b"cuda-sm100a"
    assert b"sm_100a" in gpu_lib.bf_version()


def test_oracle_exports_every_symbol(oracle_lib):
    for name in header_symbols():
        assert hasattr(oracle_lib, name), name
    assert oracle_lib.bf_backend() == b"cpu-oracle"


def test_cuda_library_is_built_for_sm100a(gpu_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", L.DEFAULT_LIBRARY], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_path_fails_loudly_without_gpu(gpu_lib):
    """No CPU fallback: without a visible device bf_create must return an error, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(BeamGpuError) as e:
        coupledTotalLagNewtonRaphsonBeam(cases.cantilever(), library=gpu_lib)
    assert e.value.status == L.BF_ERR_CUDA


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        L.load_library(str(tmp_path / "libbeamgpu.so"))


def test_interpolation_table_clamp():
    t = interpolationTable([(0, (0, 0, 0)), (1, (0, -6.28318, 0))])
    np.testing.assert_allclose(t(0.25), [0, -1.570795, 0])
    np.testing.assert_allclose(t(-1), [0, 0, 0])
    np.testing.assert_allclose(t(7), [0, -6.28318, 0])
    t = interpolationTable([(0, (8, 0, 0)), (2.5, (8, 0, 0)), (2.51, (0, 0, 0)), (1000, (0, 0, 0))])
    np.testing.assert_allclose(t(2.505), [4, 0, 0])


def test_shard_range_partitions():
    for B, G in ((65536, 8), (7, 3), (3, 8 - 5), (10, 4)):
        r = [shard_range(B, k, G) for k in range(G)]
        assert r[0][0] == 0 and r[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_argument_checking(oracle_lib):
    import dataclasses
    with pytest.raises(ValueError):  # transient scheme without density: Evolve.C:323-335
        coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(cases.cantilever(), d2dt2Scheme="Newmark"),
                                         library=oracle_lib)
    with pytest.raises(ValueError):  # CTL.C:708-720
        coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(cases.cantilever(), d2dt2Scheme="CrankNicolson"),
                                         library=oracle_lib)
    m = coupledTotalLagNewtonRaphsonBeam(dataclasses.replace(cases.dynamicCantilever(), d2dt2Scheme="Euler"),
                                         library=oracle_lib)  # all three d2dt2Schemes of CTL.C:708-720 exist
    m.close()
    m = coupledTotalLagNewtonRaphsonBeam(cases.cantilever(), library=oracle_lib)
    bad = np.array([m.nCells], dtype=np.int64)  # one past the last cell
    z = np.zeros(3)
    assert m.lib.bf_set_point_forces(m._ctx, 1, bad.ctypes.data, z.ctypes.data, z.ctypes.data) == L.BF_ERR_INVALID
    assert m.lib.bf_set_point_forces(m._ctx, 0, None, None, None) == L.BF_OK
    assert m.lib.bf_set_bc_values(m._ctx, 5, 0, None) == L.BF_ERR_INVALID
    assert m.lib.bf_begin_step(m._ctx, -1.0) == L.BF_ERR_INVALID


def test_max_iterations_and_divergence_are_status_codes(oracle_lib):
    """Evolve.C:1003-1022 abort with FatalError; across the C ABI they are BF_MAXITER / BF_DIVERGED."""
    import dataclasses
    case = dataclasses.replace(cases.largeDeflectionCantilever(0.588), nCorrectors=2)
    m = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    m.loop()
    with pytest.raises(BeamGpuError) as e:
        m.evolve()
    assert e.value.status == L.BF_MAXITER and "Maximum iterations reached" in str(e.value)
    assert m.lastStep.nIter == 3
    case = dataclasses.replace(cases.largeDeflectionCantilever(0.588), divergenceTol=1e-3)
    m = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    m.loop()
    with pytest.raises(BeamGpuError) as e:
        m.evolve()
    assert e.value.status == L.BF_DIVERGED


def test_field_roundtrip_and_ordering(oracle_lib):
    """bf_set_field/bf_get_field use createBeamMesh ordering: cells beam-major, then patch values."""
    import dataclasses
    case = dataclasses.replace(cases.multiBeamDensity(3, 5), nSegments=[4, 6, 5])
    m = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    assert m.nCells == 15 and m.nInternalFaces == 12
    rng = np.random.default_rng(0)
    for name in ("W", "Lambda", "Lambdaf", "Q"):
        nc = 9 if name.startswith("Lambda") else 3
        n = m.nInternalFaces if name in L.SURFACE_FIELDS else m.nCells
        I, lft, rgt = rng.normal(size=(n, nc)), rng.normal(size=(3, nc)), rng.normal(size=(3, nc))
        m.set_field(name, I, lft, rgt)
        I2, l2, r2 = m.get_field(name)
        np.testing.assert_array_equal(I, I2)
        np.testing.assert_array_equal(lft, l2)
        np.testing.assert_array_equal(rgt, r2)


def run_two_halves(model, nSteps):
    """The time loop of a host that evaluates the next time level's BC series while the solve of the open one runs:
    uploadPrepared / evolveBegin / updateCoeffs(t + dt, prepareOnly) / evolveEnd."""
    iters = []
    while len(iters) < nSteps and model.loop():
        if model._bcPrepared:
            model.uploadPrepared()
        else:
            model.updateCoeffs(model.time)
        model.evolveBegin()
        model.updateCoeffs((model.timeIndex + 1) * model.case.deltaT, prepareOnly=True)  # the time loop() will set
        model.evolveEnd()
        iters.append(model.iOuterCorr())
    return iters


def test_evolve_in_two_halves_is_evolve(oracle_lib):
    """bf_evolve_begin + bf_evolve_end = bf_evolve, result for result (time-dependent follower load, Evolve.C:156-157)."""
    case = cases.cantileverFollowerForce()
    a = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    b = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    assert a.run(nSteps=12) == run_two_halves(b, 12)
    for x, y in zip(a.tip(1), b.tip(1)):
        assert np.array_equal(x, y)
    # the halves come in pairs
    out = L.bf_step_out()
    assert oracle_lib.bf_evolve_end(b._ctx, C.byref(out)) == L.BF_ERR_INVALID
    assert oracle_lib.bf_evolve_begin(b._ctx, C.byref(b.tol)) == L.BF_OK
    assert oracle_lib.bf_evolve_begin(b._ctx, C.byref(b.tol)) == L.BF_ERR_INVALID
    assert oracle_lib.bf_evolve_end(b._ctx, C.byref(out)) in (L.BF_OK, L.BF_MAXITER, L.BF_DIVERGED)
    a.close()
    b.close()
