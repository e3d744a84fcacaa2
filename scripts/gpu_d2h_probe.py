"""Does a 629 MB D2H copy slow down while the Newton kernels saturate HBM, or do the kernels slow down?  (GPU box)"""
import ctypes as C, sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, _lib as L
case = cases.multiBeamDensity(65536, 200, jitter=True)
m = coupledTotalLagNewtonRaphsonBeam(case, storeIncrements=False)
lib, ctx = m.lib, m._ctx
def evolve():
    m.loop(); lib.bf_begin_step(ctx, case.deltaT); out = L.bf_step_out(); lib.bf_evolve(ctx, C.byref(m.tol), C.byref(out))
m.updateCoeffs(case.deltaT)
for _ in range(3): evolve()
n = 13107200 * 6
src = torch.zeros(n, dtype=torch.float64, device="cuda")
dst = torch.empty(n, dtype=torch.float64, pin_memory=True)
side = torch.cuda.Stream()
def copy_ms(concurrent, chunks=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        e0.record()
        step = n // chunks
        for k in range(chunks):
            dst[k * step:(k + 1) * step].copy_(src[k * step:(k + 1) * step], non_blocking=True)
        e1.record()
    t = time.perf_counter()
    if concurrent:
        evolve()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t) * 1e3
    return e0.elapsed_time(e1), wall
for conc in (False, True, True):
    for chunks in (1, 8):
        c, w = copy_ms(conc, chunks)
        print(f"concurrent={conc} chunks={chunks}: copy {c:.2f} ms ({n*8/c/1e6:.1f} GB/s), evolve+sync wall {w:.2f} ms")
