#!/usr/bin/env python
"""Measured GPU-vs-oracle tip differences of the conditioning-critical cases (to set the test tolerances)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_oracle import build as build_oracle
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
lib = load_library(); olib = load_library(build_oracle())
def probe(name, case, steps):
    g = coupledTotalLagNewtonRaphsonBeam(case, library=lib); o = coupledTotalLagNewtonRaphsonBeam(case, library=olib)
    worst = 0.0; its = []
    for k in range(steps):
        g.loop(); o.loop(); g.evolve(); o.evolve(); its.append((g.iOuterCorr(), o.iOuterCorr()))
        for a, b in zip(g.tip(1), o.tip(1)):
            if np.max(np.abs(b)) > 1e-13: worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(b))))
    fw = max(float(np.max(np.abs(g.get_field(f)[0] - o.get_field(f)[0])) / max(np.max(np.abs(o.get_field(f)[0])), 1e-300)) for f in ("W", "Theta"))
    print(f"{name:28s} steps {steps:3d} tip rel {worst:.2e} field rel {fw:.2e} iters equal {all(a == b for a, b in its)} {its[:4]}", flush=True)
    g.close(); o.close()
for F in (0.0, 0.294, 0.588):
    probe(f"largeDeflection F={F}", cases.largeDeflectionCantilever(F), 1)
probe("follower 2000 cells", cases.cantileverFollowerForce(nSegments=2000), 5)
probe("follower 10000 cells", cases.cantileverFollowerForce(nSegments=10000), 3)
probe("freeFlexibleBeam", cases.freeFlexibleBeam(), 300)
probe("cantileverSpring", cases.cantileverSpring(), 1)
