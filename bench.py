#!/usr/bin/env python
"""bench.py — beam-cell Newton updates/s on the synthetic multi-beam bundle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): 65,536 beams x 200 cells PER GPU of the
`3DdynamicCantileverMultiBeamDensity` tutorial scaled up — Newmark (beta .25, gamma .5),
dt = 1e-3, square 0.2 x 0.2 section, rho cycling 500/1000/2000, clamped left end, constant
tip force (2000 2000 0) with a deterministic +-10 % per-beam jitter, tolerances of the
tutorial (nCorrectors 200, solutionTol 1e-10 => rtol 1e-8).  A "step" is one time step =
one evolve() = one Newton loop (3-4 iterations) over every cell of the bundle.

metric  = beam-cell Newton updates per second = cells x Newton iterations / time.
value   = state resident in HBM; the K timed steps are bracketed by a barrier +
          synchronize and timed with CUDA events on the library's stream (max over ranks).
e2e     = the same K steps through the host-facing call sequence with HOST buffers inside
          the timed region: BC values host->device every step, W and Theta mirrored
          device->host into pinned buffers every step.
roofline= one Newton iteration = two launches (beam_forward, beam_backward): 760 B (Newmark)
          algorithmic bytes per beam-cell update / measured time of the pair,
          against the measured HBM copy bandwidth of MEASURED_PEAKS.json; `launches` lists each
          kernel with its own time and measured DRAM bytes.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
BYTES_PER_UPDATE_NEWMARK = 760.0  # 95 doubles
HBM_FALLBACK_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--beams", type=int, default=65536, help="beams per GPU (weak scaling)")
    ap.add_argument("--cells", type=int, default=200)
    ap.add_argument("--strong", action="store_true", help="keep the total at --beams (strong scaling)")
    ap.add_argument("--no-jitter", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--write-interval", type=int, default=10,
                    help="e2e: the full W/Theta fields go to the host every this many steps (writeInterval 10 in "
                         "tutorials/3DdynamicCantileverMultiBeamDensity/system/controlDict); the patch values every step")
    ap.add_argument("--cpu-beams-per-thread", type=int, default=256)
    return ap.parse_args()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a background thread every 5 ms (so that a
    timed region of a few steps still gets samples), nvidia-smi -lms 200 as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []   # (sm_mhz, reason bitmask)
        self.running = False
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

    def _poll(self, nv, h):
        while self.running:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.nvml = (nv, h)
            self.running = True
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            nv, h = self.nvml
            self.running = False
            self.thread.join(timeout=1.0)
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            sm = [a for a, _ in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or f[0] != str(self.gpu):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200"}


def make_case(args, world):
    from beamfoam_b200 import cases
    total = args.beams if args.strong else args.beams * world
    return cases.multiBeamDensity(total, args.cells, jitter=not args.no_jitter), total


def config_dict(args, world, total):
    return {
        "workload": f"3DdynamicCantileverMultiBeamDensity scaled: {total} beams x {args.cells} cells total "
                    f"({total // world} beams per GPU), Newmark beta .25 gamma .5 dt 1e-3, "
                    f"rho 500/1000/2000, tip force (2000 2000 0)" + ("" if args.no_jitter else " x (1+0.1u_b)"),
        "beams_total": total, "beams_per_gpu": total // world, "cells_per_beam": args.cells,
        "cells_total": total * args.cells, "tolerances": "nCorrectors 200, rtol 1e-8, stol 1e-10",
        "l2_policy": f"inputs larger than L2: {total // world * args.cells * 280 / 1e9:.1f} GB of state + {total // world * args.cells * 336 / 1e9:.1f} GB of scratch per GPU vs 126 MB L2",
        "parallelism": f"beams sharded over {world} GPU(s); one 3-double all-reduce per Newton iteration",
    }


# ----------------------------------------------------------------------------- CPU arms
def cpu_oracle_rate(args, steps, warmup, threads=None):
    """Times the CPU oracle (a port of the reference arithmetic; the reference binary cannot be
    built here) on a bounded sample of the same workload: `threads` independent shards of
    `cpu_beams_per_thread` beams x cells, one oracle context per host thread (ctypes releases
    the GIL).  Returns (updates/s, description, threads)."""
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from build_oracle import build as build_oracle
    from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
    lib = load_library(build_oracle())
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
    _, total = make_case_sizes(args, world)
    rate, desc, threads, ms = cpu_oracle_rate(args, args.steps, args.warmup)
    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world, total), "impl": "reference",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle port of the reference arithmetic (OpenFOAM+Eigen cannot be built here)",
    }
    print(json.dumps(line), flush=True)


def make_case_sizes(args, world):
    total = args.beams if args.strong else args.beams * world
    return None, total


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from beamfoam_b200 import _lib as L
    from beamfoam_b200 import coupledTotalLagNewtonRaphsonBeam, load_library, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = load_library()
    case, total = make_case(args, world)
    lo, hi = shard_range(total, rank, world)
    model = coupledTotalLagNewtonRaphsonBeam(case, library=lib, device=local, beamRange=(lo, hi),
                                             storeIncrements=False)
    ctx = model._ctx
    if world > 1:  # built-in reducer: ncclAllReduce of the 3 squared norms on the library's stream
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            rc = lib.bf_nccl_unique_id(buf)
            if rc:
                raise SystemExit("bf_nccl_unique_id failed")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = (C.c_char * 128).from_buffer_copy(uid.cpu().numpy().tobytes())
        model._check(lib.bf_comm_init_nccl(ctx, raw, world, rank))

    nCellsLocal = model.nCells

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def step_resident():
        model.loop()
        model._check(lib.bf_begin_step(ctx, float(case.deltaT)))
        out = L.bf_step_out()
        rc = lib.bf_evolve(ctx, C.byref(model.tol), C.byref(out))
        model._check(rc)
        return out.nIter

    # ---- warm-up (BC values uploaded once: they are constant in this workload)
    model.updateCoeffs(case.deltaT)
    for _ in range(args.warmup):
        step_resident()

    # ---- timed region 1: state resident in HBM
    sampler = ClockSampler(local)
    lib.bf_set_timing(ctx, 1)
    lib.bf_kernel_time_ms(ctx, 1, None, None, None)
    launches0 = lib.bf_launch_count(ctx)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    lib.bf_region_mark(ctx, 0)
    iters = 0
    for _ in range(args.steps):
        iters += step_resident()
    lib.bf_region_mark(ctx, 1)
    ms = C.c_double()
    model._check(lib.bf_region_elapsed_ms(ctx, C.byref(ms)))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    launches = lib.bf_launch_count(ctx) - launches0
    fwd, bwd, nit = C.c_double(), C.c_double(), C.c_int64()
    lib.bf_kernel_time_ms(ctx, 1, C.byref(fwd), C.byref(bwd), C.byref(nit))
    lib.bf_set_timing(ctx, 0)
    region_ms = max_over_ranks(ms.value)
    updates_total = sum_over_ranks(float(nCellsLocal) * iters)
    value = updates_total / (region_ms * 1e-3)

    # ---- roofline of one Newton iteration (this rank): two launches, beam_forward then beam_backward; the algorithmic
    # bytes of the contract (every state array read once and written once per iteration) belong to the pair
    kern_ms = (fwd.value + bwd.value) / max(nit.value, 1)
    peak, peak_src = hbm_peak()
    achieved = BYTES_PER_UPDATE_NEWMARK * nCellsLocal / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    traffic, per_kernel, fp64_per_update = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("newton_iteration_dram_bytes")
        scale = nCellsLocal / float(tj.get("cells_per_launch", nCellsLocal))
        fms, bms = fwd.value / max(nit.value, 1), bwd.value / max(nit.value, 1)
        per_kernel = [
            {"kernel": "beam_forward<BFK_NEWMARK>", "ms": fms, "dram_bytes": tj.get("beam_forward_dram_bytes"),
             "dram_GBs": tj.get("beam_forward_dram_bytes", 0) * scale / (fms * 1e-3) / 1e9 if fms > 0 else None,
             "fp64_TFLOPs": tj.get("beam_forward_fp64_flop", 0) * scale / (fms * 1e-3) / 1e12 if fms > 0 else None},
            {"kernel": "beam_backward<BFK_NEWMARK>", "ms": bms, "dram_bytes": tj.get("beam_backward_dram_bytes"),
             "dram_GBs": tj.get("beam_backward_dram_bytes", 0) * scale / (bms * 1e-3) / 1e9 if bms > 0 else None,
             "fp64_TFLOPs": tj.get("beam_backward_fp64_flop", 0) * scale / (bms * 1e-3) / 1e12 if bms > 0 else None}]
        fp64_per_update = tj.get("fp64_flop_per_update")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "beam_forward<BFK_NEWMARK> + beam_backward<BFK_NEWMARK> (one Newton iteration)",
                "kernel_ms": kern_ms, "bytes_per_update": BYTES_PER_UPDATE_NEWMARK, "updates_per_launch": nCellsLocal,
                "peak_source": peak_src, "kernel_share_of_step": (fwd.value + bwd.value) / ms.value,
                "fwd_ms": fwd.value / max(nit.value, 1), "bwd_ms": bwd.value / max(nit.value, 1),
                "launches": per_kernel,
                "fp64": {"flop_per_update": fp64_per_update,
                         "achieved_TFLOPs": fp64_per_update * nCellsLocal / (kern_ms * 1e-3) / 1e12 if fp64_per_update and kern_ms > 0 else None,
                         "note": "thread-level DFMA x2 + DMUL + DADD of one ncu capture (profiles/traffic.json); B200 FP64 peak ~37 TFLOP/s"}}

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

        tipW = np.empty((B, 3))
        tipT = np.empty((B, 3))

        def step_e2e(k, interval):
            """One step as the reference's beamFoam.C:73-89 runs it: BC series -> device, evolve(), the function objects'
            per-step reads (patch W, Theta of every beam: functionObjects/beamDisplacements), and at write times
            (every `interval` steps, runTime.write()) the full W, Theta fields."""
            model.loop()
            model.updateCoeffs(model.time)  # host -> device
            model._check(lib.bf_begin_step(ctx, float(case.deltaT)))
            out = L.bf_step_out()
            model._check(lib.bf_evolve(ctx, C.byref(model.tol), C.byref(out)))
            model._check(lib.bf_get_field(ctx, wf, None, None, tipW.ctypes.data))
            model._check(lib.bf_get_field(ctx, tf, None, None, tipT.ctypes.data))
            if (k + 1) % interval == 0:
                # queued on the copy stream: overlaps the next step's Newton loop
                model._check(lib.bf_mirror_begin(ctx, 2, fields, hosts))
            return out.nIter

        def timed_e2e(interval, steps):
            step_e2e(interval - 1, interval)
            model._check(lib.bf_mirror_wait(ctx))
            h2d0 = model.h2dBytes
            barrier()
            lib.bf_region_mark(ctx, 0)
            t0 = time.perf_counter()
            it2 = 0
            for k in range(steps):
                it2 += step_e2e(k, interval)
            model._check(lib.bf_mirror_wait(ctx))  # the last write's fields must be on the host too
            lib.bf_region_mark(ctx, 1)
            ms2 = C.c_double()
            model._check(lib.bf_region_elapsed_ms(ctx, C.byref(ms2)))
            barrier()
            wall2 = (time.perf_counter() - t0) * 1e3
            e2e_ms = max_over_ranks(max(ms2.value, wall2))
            return sum_over_ranks(float(nCellsLocal) * it2) / (e2e_ms * 1e-3), e2e_ms / steps, (model.h2dBytes - h2d0) // steps

        iv = max(1, args.write_interval)
        steps_iv = max(iv, (args.steps // iv) * iv)  # whole write intervals
        rate_iv, ms_iv, h2d = timed_e2e(iv, steps_iv)
        rate_1, ms_1, _ = timed_e2e(1, args.steps) if iv != 1 else (rate_iv, ms_iv, h2d)
        tips = 2 * B * 3 * 8
        e2e = {"value": rate_iv, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": tips + d2h // iv, "ms_per_step": ms_iv, "steps": steps_iv, "write_interval": iv,
               "mirrored": f"every step: the time-series BC values (force, moment of every beam) -> device, patch W and Theta of every beam -> host; every {iv} steps "
                           "(the tutorial's writeInterval) the W, Theta internal fields -> pinned host, overlapped with "
                           "the next Newton loop, the last copy waited for inside the timed region",
               "full_mirror_every_step": {"value": rate_1, "ms_per_step": ms_1, "d2h_bytes_per_step": tips + d2h,
                                          "note": "W, Theta internal fields -> host after EVERY step (PCIe-bound)"},
               "checksum": float(Wh[-1, 0]) + float(Th[-1, 1]) + float(tipW[-1, 0])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, desc, threads, _ = cpu_oracle_rate(args, steps=8, warmup=1)
        rate1, desc1, _, _ = cpu_oracle_rate(args, steps=4, warmup=1, threads=1)  # the reference itself is serial
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
               "single_thread": {"value": rate1, "cores": 1, "sample": desc1}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": region_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, world, total), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "newton_iterations_per_step": iters / args.steps, "wall_ms_per_step": wall_ms / args.steps,
            "impl": "ours",
        }
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
