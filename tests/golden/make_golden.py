#!/usr/bin/env python
"""Generates tests/golden/oracle_*.npz from the CPU oracle (run in the build container):

    python tests/golden/make_golden.py

Each file holds, for one tutorial replica: per-step Newton iteration counts, per-step tip
(right patch) W and Theta of every beam, and the final W/Theta/Q/M fields.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from build_oracle import build as build_oracle  # noqa: E402
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library  # noqa: E402

# name -> (case, number of steps)
GOLDEN = {
    "cantilever": (lambda: cases.cantilever(), 100),
    "cantileverFollowerForce": (lambda: cases.cantileverFollowerForce(), 100),
    "largeDeflectionCantilever": (lambda: cases.largeDeflectionCantilever(0.588), 1),
    "3dDynamicCantilever": (lambda: cases.dynamicCantilever(), 100),
    "3DdynamicCantileverMultiBeamDensity": (lambda: cases.multiBeamDensity(3, 40), 60),
    "freeFlexibleBeam": (lambda: cases.freeFlexibleBeam(), 100),
    "beamAndSpringInSeries": (lambda: cases.beamAndSpringInSeries(), 10),
    "cantileverSpring": (lambda: cases.cantileverSpring(), 1),
}


def generate(name, lib):
    mk, nSteps = GOLDEN[name]
    m = coupledTotalLagNewtonRaphsonBeam(mk(), library=lib)
    iters, tipW, tipT = [], [], []
    for _ in range(nSteps):
        assert m.loop()
        m.evolve()
        iters.append(m.iOuterCorr())
        w, t = m.tip(1)
        tipW.append(w.copy())
        tipT.append(t.copy())
    out = dict(iters=np.array(iters), tipW=np.array(tipW), tipTheta=np.array(tipT))
    for f in ("W", "Theta", "Q", "M"):
        out[f] = m.get_field(f)[0]
    m.close()
    return out


if __name__ == "__main__":
    lib = load_library(build_oracle())
    for name in GOLDEN:
        out = generate(name, lib)
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **out)
        print(name, "steps", len(out["iters"]), "iterations", out["iters"][:6], "...")
