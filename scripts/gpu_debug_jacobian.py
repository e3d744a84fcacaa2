#!/usr/bin/env python
"""Debug: one Newton iteration of the same state through two kernel families; compares the increments DW / DTheta."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_oracle import build as build_oracle
from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
lib = load_library(); olib = load_library(build_oracle())
case = cases.cantileverFollowerForce()
def mk(env, library=lib):
    for k in ("BEAMGPU_LONG_BEAMS", "BEAMGPU_COOP", "BEAMGPU_SPLIT"): os.environ.pop(k, None)
    os.environ.update(env)
    return coupledTotalLagNewtonRaphsonBeam(case, library=library, storeIncrements=True)
ms = {"per_beam": mk({"BEAMGPU_COOP": "0", "BEAMGPU_SPLIT": "0"}), "cell": mk({"BEAMGPU_LONG_BEAMS": "1"}), "oracle": mk({}, olib)}
for step in range(2):
    for m in ms.values():
        m.loop(); m.updateCoeffs(m.time); m._check(m.lib.bf_begin_step(m._ctx, float(case.deltaT)))
    for k in range(4):
        out = {}
        for nm, m in ms.items():
            s = m.newton_iteration()
            out[nm] = (np.concatenate([m.get_field("DW")[0], m.get_field("DTheta")[0]], axis=1), s)
        ref = out["oracle"][0]
        for nm in ("per_beam", "cell"):
            d = out[nm][0]
            print(step, k, nm, "max|dx - dx_oracle| / max|dx_oracle| = %.3e" % (np.max(np.abs(d - ref)) / np.max(np.abs(ref))), "sums", out[nm][1], out["oracle"][1])
m = ms["per_beam"]
for k in range(3):
    s, eta = (C.c_double * 3)(), C.c_double()
    m._check(lib.bf_newton_iteration_checked(m._ctx, s, C.byref(eta))); print("eta", eta.value, list(s))
