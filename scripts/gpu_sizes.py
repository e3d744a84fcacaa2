#!/usr/bin/env python
"""Runs on the GPU box: per-iteration forward / backward kernel times at several bundle sizes, for the one-thread-per-beam
kernels (BEAMGPU_COOP=0) and the group-cooperative kernels (BEAMGPU_COOP=1).  usage: gpu_sizes.py beams [beams ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfoam_b200 import _lib as L, cases, coupledTotalLagNewtonRaphsonBeam, load_library  # noqa: E402

lib = load_library()
cells = int(os.environ.get("CELLS", "200"))
for nb in map(int, sys.argv[1:]):
    case = cases.multiBeamDensity(nb, cells, jitter=True)
    for mode in os.environ.get("MODES", "0,1,s,sc").split(","):  # 0 per beam, 1 group (warp-specialised), s split, sc split + cell-parallel backward
        # w16 / w32: split + warp-specialised forward (16 / 32 half beams per block) + cell-parallel backward; w16b / w32b: with
        # the split backward sweep; g32: unsplit, 32 beams per block
        os.environ["BEAMGPU_COOP"] = {"s": "0", "sc": "bwd", "w16": "bwd", "w32": "bwd", "w33": "bwd", "w16b": "0", "w32b": "0", "w33b": "0", "g32": "1", "g33": "1"}.get(mode, mode)
        os.environ["BEAMGPU_SPLIT"] = "1" if mode in ("s", "sc", "w16", "w32", "w33", "w16b", "w32b", "w33b") else "0"
        os.environ.pop("BEAMGPU_WS", None)
        if mode[0] in "wg":
            os.environ["BEAMGPU_WS"] = mode[1:3]
        m = coupledTotalLagNewtonRaphsonBeam(case, library=lib, storeIncrements=False)
        m.updateCoeffs(case.deltaT)
        its = []
        for k in range(6):
            if k == 2:
                lib.bf_set_timing(m._ctx, 1)
                lib.bf_kernel_time_ms(m._ctx, 1, None, None, None)
            m.loop()
            m._check(lib.bf_begin_step(m._ctx, float(case.deltaT)))
            out = L.bf_step_out()
            m._check(lib.bf_evolve(m._ctx, C.byref(m.tol), C.byref(out)))
            its.append(out.nIter)
        f, b, n = C.c_double(), C.c_double(), C.c_int64()
        lib.bf_kernel_time_ms(m._ctx, 1, C.byref(f), C.byref(b), C.byref(n))
        fm, bm = f.value / n.value, b.value / n.value
        print(f"beams {nb:6d} coop={mode:>3}: fwd {fm:.3f} ms  bwd {bm:.3f} ms  iter {fm + bm:.3f} ms  "
              f"{nb * cells / (fm + bm) / 1e6:.3f} G updates/s  iters {its} tipW {m.tip(1)[0][-1]}", flush=True)
        m.close()
