"""Where does the e2e step go?  (run on the GPU box)"""
import ctypes as C, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, _lib as L
case = cases.multiBeamDensity(65536, 200, jitter=True)
m = coupledTotalLagNewtonRaphsonBeam(case, storeIncrements=False)
lib, ctx = m.lib, m._ctx
Wh = torch.empty((m.nCells, 3), dtype=torch.float64, pin_memory=True)
Th = torch.empty((m.nCells, 3), dtype=torch.float64, pin_memory=True)
fields = (C.c_int32 * 2)(0, 1); hosts = (C.c_void_p * 2)(Wh.data_ptr(), Th.data_ptr())
def T(f, n=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
def evolve():
    m.loop(); lib.bf_begin_step(ctx, case.deltaT); out = L.bf_step_out(); lib.bf_evolve(ctx, C.byref(m.tol), C.byref(out))
m.updateCoeffs(case.deltaT)
for _ in range(3): evolve()
print("evolve only        %.2f ms" % T(evolve))
print("updateCoeffs       %.2f ms" % T(lambda: m.updateCoeffs(m.time)))
print("bc_values (numpy)  %.2f ms" % T(lambda: [m._bc_values(e, v, m.time) for e in (0, 1) for v in ("W", "Theta")]))
def mirror_sync():
    lib.bf_mirror_begin(ctx, 2, fields, hosts); lib.bf_mirror_wait(ctx)
print("mirror begin+wait  %.2f ms" % T(mirror_sync))
def both():
    evolve(); lib.bf_mirror_begin(ctx, 2, fields, hosts)
t = T(both, 8); lib.bf_mirror_wait(ctx)
print("evolve + async mirror  %.2f ms" % t)
def full():
    m.updateCoeffs(m.time); evolve(); lib.bf_mirror_begin(ctx, 2, fields, hosts)
t = T(full, 8); lib.bf_mirror_wait(ctx)
print("updateCoeffs + evolve + async mirror  %.2f ms" % t)
