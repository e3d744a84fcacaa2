#!/usr/bin/env python
"""bench.py — beam-cell Newton updates/s on the synthetic multi-beam bundle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

Workloads (BASELINE.json configs):
  bundle65536 (default, configs[3]): 65,536 beams x 200 cells IN TOTAL of the `3DdynamicCantileverMultiBeamDensity`
      tutorial scaled up, sharded by beam over the N GPUs (strong scaling: the north-star's split) — Newmark (beta .25,
      gamma .5), dt = 1e-3, square 0.2 x 0.2 section, rho cycling 500/1000/2000, clamped left end, constant tip force
      (2000 2000 0) with a deterministic +-10 % per-beam jitter, tolerances of the tutorial (nCorrectors 200,
      solutionTol 1e-10 => rtol 1e-8).  With N > 1 the same run also measures 65,536 beams PER GPU (weak scaling) and
      reports it under the key "weak".
  bundle4096 (configs[2]): 4096 beams x 200 cells of the same bundle on one B200.
  long (configs[1]): one cantilever of 10^4 cells under a follower tip force (steady state, load stepping).
A "step" is one time / load step = one evolve() = one Newton loop (3-4 iterations) over every cell.

metric   = beam-cell Newton updates per second = cells x Newton iterations / time.
value    = state resident in HBM; the K timed steps are bracketed by a barrier + synchronize and timed with CUDA
           events on the library's stream (max over ranks).
e2e      = the same K steps through the host-facing call sequence with HOST buffers inside the timed region: the
           time-series BC values host->device and the patch W / Theta of every beam device->host EVERY step, the W and
           Theta internal fields device->host into pinned buffers every `writeInterval` steps.
roofline = one Newton iteration = two launches (forward, backward sweep): 760 B (Newmark; 552 B steady) algorithmic
           bytes per beam-cell update / measured time of the pair, against the measured HBM copy bandwidth
           of MEASURED_PEAKS.json; `launches` lists each kernel with its own time and measured DRAM bytes.
parity   = checked INSIDE the run (a failed check aborts the bench): (1) a sample of this rank's beams is recomputed as a
           stand-alone 1-GPU bundle with the run's per-step iteration counts and must give bit-identical tips; (2) with
           N > 1, rank 0 re-runs the WHOLE bundle on its own GPU: the per-step Newton iteration counts must equal the
           sharded run's and the tips of its shard must agree to 1e-10.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "beam-cell Newton updates/s"
UNIT = "updates/s"
BYTES_PER_UPDATE = {"Newmark": 760.0, "steadyState": 552.0, "Euler": 760.0}  # 95 / 69 doubles
HBM_FALLBACK_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback
FAMILY = {0: "one thread per beam (beam_forward / beam_backward)",
          1: "assembler + eliminator warps (beam_forward_ws), serial back-substitution + one thread per cell (beam_backsub, beam_update)",
          2: "one thread per cell, block cyclic reduction (long_*)",
          3: "two threads per beam meeting in the middle (beam_forward_split / beam_backward_split)",
          4: "two threads per beam in the forward sweep (beam_forward_split), serial back-substitution + one thread per cell "
             "(beam_backsub_split, beam_update)",
          5: "assembler + eliminator warps per half beam (beam_forward_ws, split), serial back-substitution + one thread per cell "
             "(beam_backsub_split, beam_update)", -1: "cpu"}
# BEAMGPU_COOP, BEAMGPU_SPLIT, BEAMGPU_WS that force a family
FAMILY_ENV = {0: ("0", "0", None), 1: ("1", "0", None), 3: ("0", "1", None), 4: ("bwd", "1", None), 5: ("bwd", "1", "16")}
CONFIGS = {
    "bundle65536": dict(beams=65536, cells=200, baseline="configs[3]"),
    "bundle4096": dict(beams=4096, cells=200, baseline="configs[2]"),
    "long": dict(beams=1, cells=10000, baseline="configs[1]"),
}
KERNEL_SOURCES = ["beam_kernels.cuh", "beam_coop.cuh", "beam_long.cuh", "beam_math.cuh"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bundle65536", choices=sorted(CONFIGS))
    ap.add_argument("--beams", type=int, default=None, help="beams of the bundle IN TOTAL (default: the config's)")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--weak", action="store_true", help="--beams per GPU instead of in total (weak scaling only)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the secondary weak-scaling measurement")
    ap.add_argument("--no-verify", action="store_true", help="skip the in-run parity checks")
    ap.add_argument("--no-jitter", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--write-interval", type=int, default=10,
                    help="e2e: the full W/Theta fields go to the host every this many steps (writeInterval 10 in "
                         "tutorials/3DdynamicCantileverMultiBeamDensity/system/controlDict); the patch values every step")
    ap.add_argument("--cpu-beams-per-thread", type=int, default=256)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.beams is None:
        args.beams = cfg["beams"]
    if args.cells is None:
        args.cells = cfg["cells"]
    return args


def kernel_sources_sha():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "beamfoam_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML, on request: the timed loop calls kick()
    before a step, a waiting thread wakes, lets the step's launches go out (0.3 ms) and reads the clock and the event
    reasons of this rank's GPU ONCE, while the GPU is busy with that step and the main thread sits in its stream
    synchronisation.  Free-running pollers are not neutral when 8 ranks advance in lockstep: a 5 ms polling thread per rank
    (interpreter lock + the driver's locks of the process that is launching kernels) delayed some rank in nearly every
    3 ms step, which every other rank then waits for in the next convergence test (+0.18 ms per step, 6 %); polling child
    processes stalled steps for milliseconds.  Every rank samples its own GPU; gather() combines the ranks."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (sm_mhz, reason bitmask)
        self.req = threading.Event()
        self.running = False
        self.nvml = None
        self.thread = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # the CUDA ordinal of this process and the NVML index agree unless CUDA_VISIBLE_DEVICES reorders: go by UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def _serve(self, nv, h):
        while True:
            self.req.wait()
            self.req.clear()
            if not self.running:
                return
            time.sleep(0.0003)
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass

    def start(self):
        try:
            nv, h = self._nvml_handle()
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)  # the first call is the slow one: outside the timed region
            self.nvml = (nv, h)
            self.running = True
            self.thread = threading.Thread(target=self._serve, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def kick(self):
        if self.nvml is not None:
            self.req.set()

    def stop(self):
        """This rank's samples: {"sm": [...], "reasons": bitmask-or, "max": MHz} (None without NVML)."""
        if self.nvml is None:
            return None
        nv, h = self.nvml
        self.running = False
        self.req.set()
        self.thread.join(timeout=1.0)
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        return {"sm": [a for a, _ in self.samples], "max": mx,
                "reasons": sorted(k for k, bit in bits.items() if any(r & bit for _, r in self.samples))}

    @staticmethod
    def gather(D, mine):
        """The JSON line's `clocks` from every rank's samples (rank 0 reports; the collective runs outside the timed region)."""
        allr = [mine]
        if D.world > 1:
            import torch.distributed as dist
            allr = [None] * D.world
            dist.all_gather_object(allr, mine)
        allr = [r for r in allr if r]
        if not allr:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        sm = [x for r in allr for x in r["sm"]]
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r["max"] or 0.0) for r in allr) or None,
               "reasons": sorted({x for r in allr for x in r["reasons"]}), "samples": len(sm),
               "source": "nvml, one reading per rank and sampled step, taken while the step runs on the GPU"}
        if D.world > 1:
            out["sm_mhz_slowest_gpu"] = min(statistics.median(r["sm"]) for r in allr if r["sm"]) if sm else None
        return out


def make_case(args, total):
    from beamfoam_b200 import cases
    if args.config == "long":
        return cases.cantileverFollowerForce(nSegments=args.cells)
    return cases.multiBeamDensity(total, args.cells, jitter=not args.no_jitter)


def config_dict(args, world, total, scaling):
    per = -(-total // world)
    if args.config == "long":
        wl = (f"tutorials/cantileverFollowerForce with {args.cells} cells (BASELINE configs[1]): one cantilever, follower tip "
              f"force ramp to (0 0 1.34e5), steady state, load step 1e-3")
        l2 = ("the whole problem (14 MB of state and scratch) fits the 126 MB L2: L2 is flushed between timed steps "
              "(a 256 MB buffer is overwritten), every step is timed on its own and the times are summed")
    else:
        wl = (f"3DdynamicCantileverMultiBeamDensity scaled ({CONFIGS[args.config]['baseline'] if total == CONFIGS[args.config]['beams'] else 'custom size'}): "
              f"{total} beams x {args.cells} cells in total ({per} beams per GPU), Newmark beta .25 gamma .5 dt 1e-3, "
              f"rho 500/1000/2000, tip force (2000 2000 0)" + ("" if args.no_jitter else " x (1+0.1u_b)"))
        l2 = (f"inputs larger than L2: {per * args.cells * 280 / 1e9:.2f} GB of state + {per * args.cells * 336 / 1e9:.2f} GB "
              f"of scratch per GPU vs 126 MB L2")
    return {
        "workload": wl, "config": args.config, "beams_total": total, "beams_per_gpu": per, "cells_per_beam": args.cells,
        "cells_total": total * args.cells, "tolerances": "the tutorial's (nCorrectors, residualTol, solutionTol unchanged)",
        "l2_policy": l2,
        "parallelism": f"beams sharded over {world} GPU(s) ({scaling} scaling); the only exchange is the sum of 3 doubles per Newton iteration",
    }


# ----------------------------------------------------------------------------- CPU arms
def cpu_oracle_rate(args, steps, warmup, threads=None):
    """Times the CPU oracle (a port of the reference arithmetic; the reference binary cannot be
    built here) on a bounded sample of the same workload: `threads` independent shards of
    `cpu_beams_per_thread` beams x cells, one oracle context per host thread (ctypes releases
    the GIL).  Returns (updates/s, description, threads, ms per step)."""
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from build_oracle import build as build_oracle
    from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
    lib = load_library(build_oracle())
    if args.config == "long":  # one beam: one chain, one thread (the reference is serial as well)
        threads, bpt = 1, 1
        case = make_case(args, 1)
        models = [coupledTotalLagNewtonRaphsonBeam(case, library=lib, storeIncrements=False)]
    else:
        threads = threads or os.cpu_count() or 1
        bpt = args.cpu_beams_per_thread
        case = cases.multiBeamDensity(threads * bpt, args.cells, jitter=not args.no_jitter)
        models = [coupledTotalLagNewtonRaphsonBeam(case, library=lib, beamRange=(t * bpt, (t + 1) * bpt),
                                                   storeIncrements=False) for t in range(threads)]

    def work(m, n):
        it = 0
        for _ in range(n):
            m.loop()
            m.evolve()
            it += m.iOuterCorr()
        return it

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda m: work(m, warmup), models))
        t0 = time.perf_counter()
        iters = list(ex.map(lambda m: work(m, steps), models))
        dt = time.perf_counter() - t0
    updates = sum(iters) * bpt * args.cells
    for m in models:
        m.close()
    desc = (f"{threads} threads x {bpt} beams x {args.cells} cells, {steps} time steps "
            f"({sum(iters) // threads} Newton iterations per shard), oracle port, banded LU per beam")
    return updates / dt, desc, threads, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    total = args.beams * world if args.weak else args.beams
    rate, desc, threads, ms = cpu_oracle_rate(args, args.steps, args.warmup)
    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world, total, "weak" if args.weak else "strong"), "impl": "reference",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle port of the reference arithmetic (OpenFOAM+Eigen cannot be built here); the rate of "
                "a bounded sample of the workload on the host cores of ONE box, independent of --gpus",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class Dist:
    """rank / world plumbing (torch.distributed over NCCL when launched by torchrun)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def min(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def nccl_id(self, lib):
        """ncclUniqueId of a new communicator for the library's own all-reduce, made on rank 0."""
        torch = self.torch
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            if lib.bf_nccl_unique_id(buf):
                raise SystemExit("bf_nccl_unique_id failed")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        self.dist.broadcast(uid, 0)
        return (C.c_char * 128).from_buffer_copy(uid.cpu().numpy().tobytes())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


class Bundle:
    """This rank's shard of a bundle behind the reference-facing host object, plus the step functions of the bench."""

    def __init__(self, args, D: Dist, lib, case, total, beamRange=None, comm=True):
        from beamfoam_b200 import coupledTotalLagNewtonRaphsonBeam, shard_range
        self.args, self.D, self.lib, self.case = args, D, lib, case
        self.range = beamRange if beamRange is not None else shard_range(total, D.rank, D.world)
        self.model = coupledTotalLagNewtonRaphsonBeam(case, library=lib, device=D.local, beamRange=self.range,
                                                      storeIncrements=False)
        self.ctx = self.model._ctx
        self.exchange = "none (one GPU)"
        if comm and D.world > 1:
            # fallback reducer: ncclAllReduce of the 3 squared norms on the library's stream, one host round trip per iteration
            self.model._check(lib.bf_comm_init_nccl(self.ctx, D.nccl_id(lib), D.world, D.rank))
            self.exchange = "ncclAllReduce per iteration"
            if not os.environ.get("BEAMGPU_NO_IPC"):
                # one-sided exchange over NVLink peer memory + device-side convergence test (bf_comm_ipc_attach)
                h = (C.c_char * 64)()
                ok = lib.bf_comm_ipc_export(self.ctx, h) == 0
                mine = D.torch.frombuffer(bytearray(h.raw), dtype=D.torch.uint8).clone().cuda()
                allh = [D.torch.empty_like(mine) for _ in range(D.world)]
                D.dist.all_gather(allh, mine)
                blob = b"".join(t.cpu().numpy().tobytes() for t in allh)
                ok = ok and lib.bf_comm_ipc_attach(self.ctx, D.world, D.rank, blob) == 0
                if D.min(1.0 if ok else 0.0) == 1.0:
                    self.exchange = "one-sided peer-memory stores (CUDA IPC over NVLink), convergence test on the device"
                elif ok:
                    raise SystemExit("bench.py: peer memory mapped on some ranks only")
                else:
                    sys.stderr.write("bench.py: peer memory not available: " + lib.bf_last_error(self.ctx).decode() + "\n")
        self.cells = self.model.nCells
        self.timeDependent = args.config == "long"  # the bundle's loads are constant in time: handed over once
        self.family = int(lib.bf_kernel_family(self.ctx))
        self.model.updateCoeffs(case.deltaT)

    def begin(self):
        m = self.model
        m.loop()
        if self.timeDependent:
            m.updateCoeffs(m.time)
        m._check(self.lib.bf_begin_step(self.ctx, float(self.case.deltaT)))

    def step_resident(self):
        from beamfoam_b200 import _lib as L
        self.begin()
        out = L.bf_step_out()
        self.model._check(self.lib.bf_evolve(self.ctx, C.byref(self.model.tol), C.byref(out)))
        return out.nIter

    def step_fixed(self, nIter):
        """The same step with a prescribed number of Newton iterations (no convergence test)."""
        self.begin()
        s = (C.c_double * 3)()
        for _ in range(nIter):
            self.model._check(self.lib.bf_newton_iteration(self.ctx, s))

    def timed_resident(self, steps, warmup, flush=None):
        """W untimed + K timed steps with the state resident in HBM.  Returns a dict; times are CUDA events on the
        library's stream, max over ranks.  flush: callable run between timed steps (then every step is timed alone)."""
        lib, D, ctx = self.lib, self.D, self.ctx
        iters = [self.step_resident() for _ in range(warmup)]
        sampler = ClockSampler(D.local)
        lib.bf_set_timing(ctx, 1)
        lib.bf_kernel_time_ms(ctx, 1, None, None, None)
        launches0 = lib.bf_launch_count(ctx)
        ms = C.c_double()
        sampler.start()
        stride = max(1, steps // 10)  # about ten clock readings per rank over the region
        D.barrier()
        t0 = time.perf_counter()
        region = 0.0
        if flush is None:
            lib.bf_region_mark(ctx, 0)
            for i in range(steps):
                if i % stride == 0:
                    sampler.kick()
                iters.append(self.step_resident())
            lib.bf_region_mark(ctx, 1)
            self.model._check(lib.bf_region_elapsed_ms(ctx, C.byref(ms)))
            region = ms.value
        else:
            for i in range(steps):
                flush()
                if i % stride == 0:
                    sampler.kick()
                lib.bf_region_mark(ctx, 0)
                iters.append(self.step_resident())
                lib.bf_region_mark(ctx, 1)
                self.model._check(lib.bf_region_elapsed_ms(ctx, C.byref(ms)))
                region += ms.value
        D.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = ClockSampler.gather(D, sampler.stop())
        fwd, bwd, nit = C.c_double(), C.c_double(), C.c_int64()
        lib.bf_kernel_time_ms(ctx, 1, C.byref(fwd), C.byref(bwd), C.byref(nit))
        lib.bf_set_timing(ctx, 0)
        timed = sum(iters[warmup:])
        region_ms = D.max(region)
        return {"iters": iters, "timed_iters": timed, "region_ms": region_ms, "local_region_ms": region, "wall_ms": wall_ms,
                "value": D.sum(float(self.cells) * timed) / (region_ms * 1e-3), "clocks": clocks,
                "launches": int(lib.bf_launch_count(ctx) - launches0), "fwd_ms": fwd.value, "bwd_ms": bwd.value,
                "nit": max(int(nit.value), 1)}

    def close(self):
        self.model.close()


def verify(args, D: Dist, lib, case, total, main: Bundle, iters, tips):
    """In-run parity (see the module docstring).  Raises SystemExit on a mismatch."""
    from beamfoam_b200 import coupledTotalLagNewtonRaphsonBeam
    lo, hi = main.range
    out = {}
    # (1) a sample of this rank's beams as a stand-alone bundle, same kernel family, same per-step iteration counts
    S = min(256, hi - lo)
    saved = {k: os.environ.get(k) for k in ("BEAMGPU_COOP", "BEAMGPU_SPLIT", "BEAMGPU_WS", "BEAMGPU_SPLIT_AT")}
    if int(lib.bf_split_point(main.ctx)) > 0:  # the sample must meet in the middle where the bundle does (rounding)
        os.environ["BEAMGPU_SPLIT_AT"] = str(int(lib.bf_split_point(main.ctx)))
    if main.family in FAMILY_ENV:
        os.environ["BEAMGPU_COOP"], os.environ["BEAMGPU_SPLIT"], ws = FAMILY_ENV[main.family]
        if ws:
            os.environ["BEAMGPU_WS"] = ws
    sample = Bundle(args, D, lib, case, total, beamRange=(lo, lo + S), comm=False)
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    assert sample.family == main.family, (sample.family, main.family)
    for n in iters:
        sample.step_fixed(n)
    sW, sT = sample.model.tip(1)
    sample.close()
    same = bool(np.array_equal(sW, tips[0][:S]) and np.array_equal(sT, tips[1][:S]))
    ok = D.min(1.0 if same else 0.0) == 1.0
    out["sample_beams_per_rank"] = S
    out["sample_bit_identical_to_standalone_run"] = ok
    if not ok:
        raise SystemExit(f"bench.py: rank {D.rank}: the tips of beams {lo}..{lo + S} differ from their stand-alone recomputation")
    # (2) N > 1: the whole bundle on ONE GPU (rank 0): same per-step iteration counts, tips of rank 0's shard to 1e-10
    if D.world > 1:
        bad = 0.0
        if D.rank == 0:
            m1 = coupledTotalLagNewtonRaphsonBeam(case, library=lib, device=D.local, storeIncrements=False)
            m1.updateCoeffs(case.deltaT)
            its1 = m1.run(nSteps=len(iters))
            w1, t1 = m1.tip(1)
            m1.close()
            rel = max(float(np.max(np.abs(w1[lo:hi] - tips[0])) / np.max(np.abs(w1[lo:hi]))),
                      float(np.max(np.abs(t1[lo:hi] - tips[1])) / np.max(np.abs(t1[lo:hi]))))
            out["iterations_equal_to_1gpu_run"] = its1 == list(iters)
            out["tips_vs_1gpu_run_max_rel"] = rel
            if its1 != list(iters) or not rel <= 1e-10:
                bad = 1.0
                sys.stderr.write(f"bench.py: sharded run {list(iters)} vs 1-GPU run {its1}, tips rel {rel:.3e}\n")
        if D.max(bad) > 0:
            raise SystemExit("bench.py: the sharded run does not reproduce the 1-GPU run")
    return out


def load_traffic(kind):
    """Measured DRAM traffic / FP64 work of the kernels from the committed ncu capture.  Stale numbers (the kernel sources
    changed after the capture) are NOT reported: traffic is null and a warning goes to stderr; tests/test_bench_contract.py
    fails on a stale file."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
    except Exception:
        return None
    if tj.get("kernel_sources_sha") != kernel_sources_sha():
        sys.stderr.write("bench.py: profiles/traffic.json is older than the kernel sources: roofline.traffic omitted "
                         "(re-run scripts/gpu_profile_r02.sh on the GPU box and scripts/make_profile_summary.py r02)\n")
        return None
    return tj.get(kind)


def run_ours(args):
    from beamfoam_b200 import _lib as L
    from beamfoam_b200 import load_library
    D = Dist(args)
    torch = D.torch
    if args.config != "bundle65536" and D.world > 1:
        raise SystemExit(f"--config {args.config} is a single-GPU workload (BASELINE {CONFIGS[args.config]['baseline']})")
    lib = load_library()
    scaling = "weak" if args.weak else "strong"
    total = args.beams * D.world if args.weak else args.beams
    case = make_case(args, total)
    main = Bundle(args, D, lib, case, total)
    model, ctx = main.model, main.ctx
    exchange = main.exchange
    nCellsLocal = main.cells
    flush = None
    if args.config == "long":
        junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def flush():
            junk.add_(1)
            torch.cuda.synchronize()

    # ---- timed region 1: state resident in HBM
    r = main.timed_resident(args.steps, args.warmup, flush)
    value, region_ms = r["value"], r["region_ms"]

    # ---- in-run parity
    parity = None
    if not args.no_verify:
        parity = verify(args, D, lib, case, total, main, r["iters"], model.tip(1))

    # ---- roofline of one Newton iteration (this rank): two launches, forward then backward sweep; the algorithmic
    # bytes of the contract (every state array read once and written once per iteration) belong to the pair
    bpu = BYTES_PER_UPDATE[case.d2dt2Scheme]
    nit = r["nit"]
    fms, bms = r["fwd_ms"] / nit, r["bwd_ms"] / nit
    kern_ms = fms + bms
    peak, peak_src = hbm_peak()
    achieved = bpu * nCellsLocal / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    names = {0: ("beam_forward", "beam_backward"), 1: ("beam_forward_ws", "beam_backsub + beam_update"),
             2: ("long_assemble + long_pcr_*", "long_update_*"), 3: ("beam_forward_split", "beam_backward_split"),
             4: ("beam_forward_split", "beam_backsub_split + beam_update"),
             5: ("beam_forward_ws", "beam_backsub_split + beam_update")}[main.family]
    tj = load_traffic({0: "per_beam", 1: "group", 2: "cell", 3: "split", 4: "split_cell", 5: "split_group"}[main.family])
    traffic, per_kernel, fp64 = None, None, None
    if tj:
        scale = nCellsLocal / float(tj["cells_per_launch"])
        traffic = (tj["forward_dram_bytes"] + tj["backward_dram_bytes"]) * scale
        per_kernel = [
            {"kernel": names[0], "ms": fms, "dram_bytes": tj["forward_dram_bytes"] * scale,
             "dram_GBs": tj["forward_dram_bytes"] * scale / (fms * 1e-3) / 1e9 if fms > 0 else None,
             "fp64_TFLOPs": tj["forward_fp64_flop"] * scale / (fms * 1e-3) / 1e12 if fms > 0 else None},
            {"kernel": names[1], "ms": bms, "dram_bytes": tj["backward_dram_bytes"] * scale,
             "dram_GBs": tj["backward_dram_bytes"] * scale / (bms * 1e-3) / 1e9 if bms > 0 else None,
             "fp64_TFLOPs": tj["backward_fp64_flop"] * scale / (bms * 1e-3) / 1e12 if bms > 0 else None}]
        fpu = (tj["forward_fp64_flop"] + tj["backward_fp64_flop"]) / float(tj["cells_per_launch"])
        fp64 = {"flop_per_update": fpu, "achieved_TFLOPs": fpu * nCellsLocal / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else None,
                "note": "thread-level DFMA x2 + DMUL + DADD of the ncu capture (profiles/traffic.json); B200 FP64 peak ~37 TFLOP/s"}
    else:
        per_kernel = [{"kernel": names[0], "ms": fms}, {"kernel": names[1], "ms": bms}]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": f"{names[0]} + {names[1]} (one Newton iteration, {case.d2dt2Scheme})",
                "kernel_family": FAMILY[main.family], "kernel_ms": kern_ms, "bytes_per_update": bpu,
                "updates_per_launch": nCellsLocal, "peak_source": peak_src,
                "kernel_share_of_step": (r["fwd_ms"] + r["bwd_ms"]) / r["local_region_ms"],
                "fwd_ms": fms, "bwd_ms": bms, "launches": per_kernel, "fp64": fp64}

    # ---- timed region 2: end to end with host buffers
    e2e = None
    if not args.no_e2e:
        Wh = torch.empty((nCellsLocal, 3), dtype=torch.float64, pin_memory=True)
        Th = torch.empty((nCellsLocal, 3), dtype=torch.float64, pin_memory=True)
        B = model._B
        d2h = 2 * nCellsLocal * 3 * 8
        wf, tf = L.FIELDS["W"], L.FIELDS["Theta"]
        fields = (C.c_int32 * 2)(wf, tf)
        hosts = (C.c_void_p * 2)(Wh.data_ptr(), Th.data_ptr())
        tipW, tipT = model.stepOutputs(1)  # filled by every evolve() before it returns (bf_set_step_outputs)

        def step_e2e(k, interval):
            """One step as the reference's beamFoam.C:73-89 runs it: BC series -> device, evolve(), the function objects'
            per-step reads (patch W, Theta of every beam: functionObjects/beamDisplacements), and at write times
            (every `interval` steps, runTime.write()) the full W, Theta fields."""
            model.loop()
            if model._bcPrepared:
                model.uploadPrepared()          # host -> device: the values of this time level, evaluated during the previous solve
            else:
                model.updateCoeffs(model.time)  # host -> device (first step: nothing was prepared)
            model.evolveBegin()                 # bf_begin_step + bf_evolve_begin: the Newton loop is enqueued, the call returns
            # while the GPU works: the BC series of the NEXT time level into the other set of page-locked staging arrays
            model.updateCoeffs((model.timeIndex + 1) * case.deltaT, prepareOnly=True)
            model.evolveEnd()                   # bf_evolve_end; patch W, Theta of every beam are in tipW, tipT now
            if (k + 1) % interval == 0:
                # queued on the copy stream: overlaps the next step's Newton loop
                model._check(lib.bf_mirror_begin(ctx, 2, fields, hosts))
            return model.lastStep.nIter

        def timed_e2e(interval, steps):
            step_e2e(interval - 1, interval)
            model._check(lib.bf_mirror_wait(ctx))
            h2d0 = model.h2dBytes
            D.barrier()
            lib.bf_region_mark(ctx, 0)
            t0 = time.perf_counter()
            it2 = 0
            for k in range(steps):
                it2 += step_e2e(k, interval)
            model._check(lib.bf_mirror_wait(ctx))  # the last write's fields must be on the host too
            lib.bf_region_mark(ctx, 1)
            ms2 = C.c_double()
            model._check(lib.bf_region_elapsed_ms(ctx, C.byref(ms2)))
            wall2 = (time.perf_counter() - t0) * 1e3
            D.barrier()
            e2e_ms = D.max(max(ms2.value, wall2))
            return D.sum(float(nCellsLocal) * it2) / (e2e_ms * 1e-3), e2e_ms / steps, (model.h2dBytes - h2d0) // steps

        iv = max(1, args.write_interval)
        steps_iv = max(iv, (args.steps // iv) * iv)  # whole write intervals
        rate_iv, ms_iv, h2d = timed_e2e(iv, steps_iv)
        rate_1, ms_1, _ = timed_e2e(1, args.steps) if iv != 1 else (rate_iv, ms_iv, h2d)
        tips = 2 * B * 3 * 8
        e2e = {"value": rate_iv, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": tips + d2h // iv, "ms_per_step": ms_iv, "steps": steps_iv, "write_interval": iv,
               "mirrored": f"every step: the time-series BC values (force, moment of every beam) -> device (evaluated on the host during the previous step's Newton loop, bf_evolve_begin / bf_evolve_end), patch W and Theta of every beam -> host; every {iv} steps "
                           "(the tutorial's writeInterval) the W, Theta internal fields -> pinned host, overlapped with "
                           "the next Newton loop, the last copy waited for inside the timed region",
               "full_mirror_every_step": {"value": rate_1, "ms_per_step": ms_1, "d2h_bytes_per_step": tips + d2h,
                                          "note": "W, Theta internal fields -> host after EVERY step (PCIe-bound)"},
               "checksum": float(Wh[-1, 0]) + float(Th[-1, 1]) + float(tipW[-1, 0])}
    main.close()

    # ---- N > 1: the weak-scaling companion (the config's bundle PER GPU), resident only
    weak = None
    if D.world > 1 and not args.weak and not args.no_weak:
        wtotal = args.beams * D.world
        wb = Bundle(args, D, lib, make_case(args, wtotal), wtotal)
        wr = wb.timed_resident(args.steps, args.warmup)
        weak = {"value": wr["value"], "unit": UNIT, "scaling": "weak", "beams_per_gpu": args.beams, "beams_total": wtotal,
                "ms_per_step": wr["region_ms"] / args.steps, "newton_iterations_per_step": wr["timed_iters"] / args.steps,
                "kernel_family": FAMILY[wb.family],
                "kernel_share_of_step": (wr["fwd_ms"] + wr["bwd_ms"]) / wr["local_region_ms"]}
        wb.close()

    cpu = None
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        rate, desc, threads, _ = cpu_oracle_rate(args, steps=8 if args.config != "long" else 4, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        if threads > 1:  # the reference itself is serial
            rate1, desc1, _, _ = cpu_oracle_rate(args, steps=4, warmup=1, threads=1)
            cpu["single_thread"] = {"value": rate1, "cores": 1, "sample": desc1}

    if D.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, D.world, total, scaling), "clocks": r["clocks"],
            "e2e": e2e, "gpu_launches": r["launches"], "roofline": roofline, "cpu_baseline": cpu,
            "newton_iterations_per_step": r["timed_iters"] / args.steps, "newton_iterations": r["iters"],
            "wall_ms_per_step": r["wall_ms"] / args.steps, "parity": parity, "weak": weak, "impl": "ours",
            "norm_exchange": exchange,
        }
        print(json.dumps(line), flush=True)
    D.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
