"""The multi-GPU path on CPU: world_size-2 `gloo`, beams sharded by contiguous range, the three
squared norms summed across ranks each Newton iteration.  The compute backend here
is the CPU oracle behind the same C ABI (test infrastructure); what is under test is the host
logic: sharding, the reducer hook of bf_evolve, and that the sharded run reproduces the
single-process run — same per-step iteration counts, per-beam results to rounding."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nBeams, nSteps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    from build_oracle import ORACLE_SO
    from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def reducer(arr):
        t = torch.from_numpy(arr)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    case = cases.multiBeamDensity(nBeams, 12, jitter=True)
    lo, hi = shard_range(nBeams, rank, world)
    m = coupledTotalLagNewtonRaphsonBeam(case, library=load_library(ORACLE_SO), beamRange=(lo, hi), reducer=reducer)
    its = m.run(nSteps=nSteps)
    w, t = m.tip(1)
    q.put((rank, lo, hi, its, w, t, m.lastStep.xNorm))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_process(oracle_lib):
    from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam
    nBeams, nSteps, world = 7, 6, 2
    ref = coupledTotalLagNewtonRaphsonBeam(cases.multiBeamDensity(nBeams, 12, jitter=True), library=oracle_lib)
    ref_its = ref.run(nSteps=nSteps)
    rw, rt = ref.tip(1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, nBeams, nSteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = 0
    for rank, lo, hi, its, w, t, xn in sorted(out):
        assert its == ref_its, (rank, its, ref_its)           # one GLOBAL convergence criterion
        np.testing.assert_allclose(w, rw[lo:hi], rtol=1e-12, atol=1e-16)
        np.testing.assert_allclose(t, rt[lo:hi], rtol=1e-12, atol=1e-16)
        assert abs(xn - ref.lastStep.xNorm) <= 1e-12 * ref.lastStep.xNorm  # norms are the global ones
        covered += hi - lo
    assert covered == nBeams
