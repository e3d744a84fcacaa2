"""Debug aid (GPU box): per-tile forward/backward sweep timings inside one pipelined launch."""
import os, sys, ctypes as C
import numpy as np
os.environ["BEAMGPU_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = coupledTotalLagNewtonRaphsonBeam(cases.multiBeamDensity(nb, 200, jitter=True), library=load_library())
m.run(nSteps=3)
nt = (nb + 127) // 128
buf = np.zeros((nt, 4), dtype=np.uint64)
m.lib.bf_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
n = m.lib.bf_debug_trace(m._ctx, buf.ctypes.data, nt)
t = buf.astype(np.int64); t0 = t[:, 0].min(); t = (t - t0) / 1e6
print("tiles", n, "launch span %.3f ms" % (t.max()))
print("forward : mean %.3f ms  (first 148: %.3f, rest: %.3f)" % ((t[:,1]-t[:,0]).mean(), (t[:148,1]-t[:148,0]).mean(), (t[148:,1]-t[148:,0]).mean() if nt > 148 else 0))
print("backward: mean %.3f ms  (first 148: %.3f, rest: %.3f)" % ((t[:,3]-t[:,2]).mean(), (t[:148,3]-t[:148,2]).mean(), (t[148:,3]-t[148:,2]).mean() if nt > 148 else 0))
print("backward start - forward end (wait slack): mean %.3f" % (t[:,2]-t[:,1]).mean())
for k in range(0, nt, max(1, nt // 16)):
    print("tile %4d: F %.3f-%.3f  B %.3f-%.3f" % (k, *t[k]))
print("last forward end %.3f, first backward start %.3f, last backward end %.3f" % (t[:,1].max(), t[:,2].min(), t[:,3].max()))
