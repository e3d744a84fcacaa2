"""GPU box: Newton iteration counts per step of configs[1] on both GPU paths against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from build_oracle import build as build_oracle
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
case = cases.cantileverFollowerForce(nSegments=n)
o = coupledTotalLagNewtonRaphsonBeam(case, library=load_library(build_oracle()))
oi = []
hist = []
for k in range(steps):
    o.loop(); o.evolve(); oi.append(o.iOuterCorr()); hist.append((o.lastStep.currentResidualNorm, o.lastStep.deltaXNorm, o.lastStep.xNorm, o.lastStep.initialResidualNorm))
print("oracle", oi)
for mode in ("0", "1"):
    os.environ["BEAMGPU_LONG_BEAMS"] = mode
    m = coupledTotalLagNewtonRaphsonBeam(case, library=load_library())
    gi = []
    for k in range(steps):
        m.loop(); m.evolve(); gi.append(m.iOuterCorr())
        if gi[-1] != oi[k]:
            s = m.lastStep
            print(f"  step {k}: gpu {gi[-1]} (res {s.currentResidualNorm:.3e} dx {s.deltaXNorm:.3e} x {s.xNorm:.3e} res0 {s.initialResidualNorm:.3e})  oracle {oi[k]} (res {hist[k][0]:.3e} dx {hist[k][1]:.3e} x {hist[k][2]:.3e} res0 {hist[k][3]:.3e})")
    print("mode", mode, gi, "tip diff", float(np.max(np.abs(m.tip(1)[0] - o.tip(1)[0]))))
    m.close()
