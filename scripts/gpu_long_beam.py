"""GPU box: time BASELINE configs[1] (cantileverFollowerForce, 1 beam x 10^4 cells) on the two launch paths."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
res = {}
for mode in ("0", "1"):
    os.environ["BEAMGPU_LONG_BEAMS"] = mode
    m = coupledTotalLagNewtonRaphsonBeam(cases.cantileverFollowerForce(nSegments=n), library=load_library())
    m.run(nSteps=2)
    t0 = time.perf_counter(); its = m.run(nSteps=steps); dt = time.perf_counter() - t0
    res[mode] = (m.tip(1)[0][0].copy(), its)
    print(f"BEAMGPU_LONG_BEAMS={mode}: {n} cells, {steps} steps, {sum(its)} Newton iterations, {dt*1e3/sum(its):.3f} ms/iteration, "
          f"{n*sum(its)/dt/1e6:.2f} M updates/s, launches {m.lib.bf_launch_count(m._ctx)}")
    m.close()
print("iteration counts equal:", res["0"][1] == res["1"][1], " tip diff:", np.max(np.abs(res["0"][0] - res["1"][0])))
