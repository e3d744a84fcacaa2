"""The two kernel families of the Newton iteration against the oracle, in every combination.

Small bundles run the kernels of beam_coop.cuh (warp-specialised forward sweep with two lanes per
beam in the elimination; serial back-substitution + one-thread-per-cell update), larger ones the one-thread-per-beam
sweeps of beam_kernels.cuh; `bf_create` picks by beam count.  Every other GPU test uses small bundles and therefore
exercises the first family; here BEAMGPU_COOP forces each family and each mixed pair (the forward sweep of one with the
backward sweep of the other: they share the uPrime/bPrime scratch) and the split sweeps (two threads per beam that
eliminate from both ends, uniform bundles of medium size) on the same cases, with the north-star bar: identical Newton
iteration counts per step, tips within 1e-10 relative.
"""
import dataclasses
import os

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam
from test_gpu_parity import run_pair

pytestmark = pytest.mark.gpu

# family -> (BEAMGPU_COOP, BEAMGPU_SPLIT[, BEAMGPU_WS]); BEAMGPU_WS = beams per block of the warp-specialised forward kernel
MODES = {"per_beam": ("0", "0"), "group": ("1", "0"), "group_fwd": ("fwd", "0"), "group_bwd": ("bwd", "0"), "split": ("0", "1"),
         "split_cell": ("bwd", "1"), "group32": ("1", "0", "32"), "split_group16": ("bwd", "1", "16"),
         "split_group32": ("bwd", "1", "32"), "split_group16_sweep": ("0", "1", "16"), "group33": ("1", "0", "33"),
         "split_group33": ("bwd", "1", "33")}  # 33 = 32 beams per block, two eliminator warps


def set_family(name):
    os.environ["BEAMGPU_COOP"], os.environ["BEAMGPU_SPLIT"] = MODES[name][:2]
    os.environ.pop("BEAMGPU_WS", None)
    if len(MODES[name]) > 2:
        os.environ["BEAMGPU_WS"] = MODES[name][2]


def clear_family():
    for k in ("BEAMGPU_COOP", "BEAMGPU_SPLIT", "BEAMGPU_WS"):
        os.environ.pop(k, None)


@pytest.fixture(params=sorted(MODES))
def coop_mode(request):
    set_family(request.param)
    yield request.param
    clear_family()


def test_newmark_bundle(gpu_lib, oracle_lib, coop_mode):
    """3DdynamicCantileverMultiBeamDensity: 70 beams (not a multiple of a block's 16 nor of a tile's 32) x 37 cells
    (not a multiple of the 2-cell batch)."""
    run_pair(cases.multiBeamDensity(70, 37, jitter=True), gpu_lib, oracle_lib, nSteps=25)


def test_steady_cantilever(gpu_lib, oracle_lib, coop_mode):
    run_pair(cases.cantilever(), gpu_lib, oracle_lib, nSteps=30)


def test_follower_force(gpu_lib, oracle_lib, coop_mode):
    run_pair(cases.cantileverFollowerForce(), gpu_lib, oracle_lib, nSteps=40)


def test_spring_and_free_ends(gpu_lib, oracle_lib, coop_mode):
    run_pair(cases.beamAndSpringInSeries(), gpu_lib, oracle_lib, nSteps=10)
    run_pair(cases.freeFlexibleBeam(), gpu_lib, oracle_lib, nSteps=40)


def test_euler_scheme(gpu_lib, oracle_lib, coop_mode):
    case = dataclasses.replace(cases.multiBeamDensity(9, 21, jitter=True), d2dt2Scheme="Euler")
    run_pair(case, gpu_lib, oracle_lib, nSteps=20, check_fields=False)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 9])
def test_short_beams(gpu_lib, oracle_lib, coop_mode, n):
    """Beams shorter than, equal to and just longer than one or two 2-cell batches; n = 1 closes both patches on one cell."""
    case = dataclasses.replace(cases.multiBeamDensity(5, n), nSegments=n)
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    assert g.run(nSteps=10) == o.run(nSteps=10)
    for a, b in zip(g.tip(1), o.tip(1)):
        assert np.max(np.abs(a - b)) <= 1e-10 * max(np.max(np.abs(b)), 1e-300)
    g.close()
    o.close()


def test_ragged_bundle(gpu_lib, oracle_lib, coop_mode):
    """Beams of different cell counts inside one block: beams run out of cells at different batches."""
    B = 19
    nSeg = np.array([3 + (7 * b) % 23 for b in range(B)], dtype=np.int32)
    case = dataclasses.replace(cases.multiBeamDensity(B, 10, jitter=True), nSegments=nSeg)
    g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
    o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
    assert g.run(nSteps=12) == o.run(nSteps=12)
    for f in ("W", "Theta"):
        a, b = g.get_field(f)[0], o.get_field(f)[0]
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b)) + 1e-13
    for a, b in zip(g.tip(1), o.tip(1)):
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    g.close()
    o.close()


def test_families_agree_at_4096_beams(gpu_lib):
    """BASELINE configs[2] (4096 beams x 200 cells, Newmark): the three kernel families give the same per-step iteration
    counts and tips equal to 1e-11 (they differ in the order of additions and, for the split sweeps, in the elimination
    order)."""
    case = cases.multiBeamDensity(4096, 200, jitter=True)
    out = {}
    families = ("per_beam", "group", "split", "split_cell", "group32", "split_group16", "split_group32", "split_group33")
    for mode in families:
        set_family(mode)
        m = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib, storeIncrements=False)
        out[mode] = (m.run(nSteps=4), m.tip(1))
        m.close()
    clear_family()
    for mode in families[1:]:
        assert out["per_beam"][0] == out[mode][0], mode
        for a, b in zip(out["per_beam"][1], out[mode][1]):
            assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b)), mode


def test_second_evolve_in_one_time_step(gpu_lib, oracle_lib):
    """evolve() called twice at the same time index (FSI outer iterations, a retry with relaxed tolerances): the Newmark
    predictor / the Euler old-time capture belong to bf_begin_step, so the second call continues from the converged
    state of the same time level instead of advancing it a second time."""
    for scheme in ("Newmark", "Euler"):
        case = dataclasses.replace(cases.multiBeamDensity(6, 14, jitter=True), d2dt2Scheme=scheme)
        g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
        o = coupledTotalLagNewtonRaphsonBeam(case, library=oracle_lib)
        for m in (g, o):
            m.run(nSteps=3)
            assert m.loop()
            m.evolve()
        w1 = g.tip(1)[0].copy()
        for m in (g, o):
            m.evolve()  # same time index: no bf_begin_step
        assert g.iOuterCorr() == o.iOuterCorr() <= 2
        for a, b in zip(g.tip(1), o.tip(1)):
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        assert np.max(np.abs(g.tip(1)[0] - w1)) <= 1e-8 * np.max(np.abs(w1))  # the converged step stays where it was
        for m in (g, o):
            m.run(nSteps=2)
        for a, b in zip(g.tip(1), o.tip(1)):
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        g.close()
        o.close()


@pytest.mark.parametrize("family", ["split", "split_cell", "split_group16", "split_group16_sweep"])
@pytest.mark.parametrize("at", [1, 9, 30, 36])
def test_unequal_split(gpu_lib, oracle_lib, family, at):
    """The two threads of a beam may meet anywhere (bf_create moves the meeting face of a multi-round bundle so that the rounds
    of resident blocks come out even, e.g. 150 + 50 cells for 65,536 beams): parts of 1, 9, 30 and 36 of 37 cells."""
    set_family(family)
    os.environ["BEAMGPU_SPLIT_AT"] = str(at)
    try:
        case = cases.multiBeamDensity(70, 37, jitter=True)
        m = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
        assert gpu_lib.bf_split_point(m._ctx) == at
        m.close()
        run_pair(case, gpu_lib, oracle_lib, nSteps=12)
    finally:
        os.environ.pop("BEAMGPU_SPLIT_AT", None)
        clear_family()


@pytest.mark.parametrize("family", ["split_cell", "split_group16", "group", "split", "per_beam"])
@pytest.mark.parametrize("switch", ["BEAMGPU_NO_SIDE_REDUCE=1", "BEAMGPU_UPDATE_WARP_BLOCKS=0", "BEAMGPU_NO_PIPELINE=1"])
def test_loop_variants_are_bit_identical(gpu_lib, family, switch):
    """How the Newton loop is driven does not touch a beam's arithmetic: the reduction in line instead of on the side stream,
    beam_update with four warps per block instead of one, the host instead of the device deciding on convergence -- every
    variant takes the same iterations and leaves bit-identical tips."""
    case = cases.multiBeamDensity(70, 37, jitter=True)
    k, v = switch.split("=")
    results = []
    for on in (False, True):
        set_family(family)
        if on:
            os.environ[k] = v
        try:
            g = coupledTotalLagNewtonRaphsonBeam(case, library=gpu_lib)
            its = g.run(nSteps=12)
            results.append((its, [t.copy() for t in g.tip(1)]))
            g.close()
        finally:
            os.environ.pop(k, None)
            clear_family()
    assert results[0][0] == results[1][0]
    for a, b in zip(results[0][1], results[1][1]):
        assert np.array_equal(a, b)
