"""Debug aid (GPU box): run one tutorial replica on the GPU library and on the oracle, print the first differences."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_oracle import build as build_oracle
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
name = sys.argv[1] if len(sys.argv) > 1 else "freeFlexibleBeam"
nSteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
case = getattr(cases, name)()
g = coupledTotalLagNewtonRaphsonBeam(case, library=load_library())
o = coupledTotalLagNewtonRaphsonBeam(case, library=load_library(build_oracle()))
for step in range(nSteps):
    g.loop(); o.loop()
    # iterate manually to compare each Newton iteration
    g.evolve(); o.evolve()
    print("step", step, "iters", g.iOuterCorr(), o.iOuterCorr(), "res", g.lastStep.currentResidualNorm, o.lastStep.currentResidualNorm)
    for f in ("W", "Theta", "Lambda", "Lambdaf", "Gamma", "K", "Q", "M", "U", "Accl", "Omega", "dotOmega"):
        try:
            a, b = g.get_field(f), o.get_field(f)
        except Exception as e:
            print(f, "n/a", e); continue
        for part, x, y in zip(("int", "left", "right"), a, b):
            if x is None or y is None: continue
            d = np.abs(np.asarray(x) - np.asarray(y))
            print(f"  {f:9s} {part:5s} maxdiff {d.max():.3e}  max|ref| {np.abs(y).max():.3e}  argmax {np.unravel_index(d.argmax(), d.shape)}")
