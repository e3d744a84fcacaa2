#!/usr/bin/env python
"""Where the per-step host overhead of the e2e loop goes (one GPU, --beams as given)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfoam_b200 import _lib as L, cases, coupledTotalLagNewtonRaphsonBeam, load_library
lib = load_library()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
case = cases.multiBeamDensity(nb, 200, jitter=True)
m = coupledTotalLagNewtonRaphsonBeam(case, library=lib, storeIncrements=False)
ctx = m._ctx
tipW, tipT = m.pinned((nb, 3)), m.pinned((nb, 3))
acc = {}
def tic(name, t0):
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
for k in range(25):
    if k == 5: acc.clear()
    t = time.perf_counter(); m.loop(); tic("loop", t)
    t = time.perf_counter(); m.updateCoeffs(m.time); tic("updateCoeffs", t)
    t = time.perf_counter(); m._check(lib.bf_begin_step(ctx, float(case.deltaT))); tic("begin_step", t)
    out = L.bf_step_out()
    t = time.perf_counter(); m._check(lib.bf_evolve(ctx, C.byref(m.tol), C.byref(out))); tic("evolve", t)
    t = time.perf_counter(); m._check(lib.bf_get_patch_pair(ctx, 1, tipW.ctypes.data, tipT.ctypes.data)); tic("patch_pair", t)
print(f"beams {nb}: per step (ms): " + ", ".join(f"{k} {v / 20 * 1e3:.3f}" for k, v in acc.items()))
