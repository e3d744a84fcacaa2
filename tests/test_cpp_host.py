"""host/beamFoamGPU (the OpenFOAM-free C++ host of the C ABI) against the Python host mirror.

A BeamCase is written out as an OpenFOAM case directory (the dictionaries the reference reads:
constant/beamProperties, constant/g, 0/W, 0/Theta, 0/q, system/controlDict, system/fvSchemes, the
(time value) tables, and an Allrun with the createBeamMesh / setInitialPositionBeam / beamFoam lines),
the C++ host runs its Allrun, and the histories it writes must equal what the Python host computes
through the same library.  CPU tests bind the oracle library (test infrastructure); the GPU test binds
libbeamgpu.so.  When the reference checkout is present its shipped tutorial directories are run UNCHANGED.
"""
import dataclasses
import os
import shutil
import subprocess

import numpy as np
import pytest

from beamfoam_b200 import cases, coupledTotalLagNewtonRaphsonBeam, load_library
from beamfoam_b200 import fvPatchFields as PF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host", "beamFoamGPU")
REF_TUTORIALS = "/root/reference/tutorials"

HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    object      {obj};
}}
"""


@pytest.fixture(scope="module")
def host_exe():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    return HOST


def _vec(v):
    return "(" + " ".join(repr(float(x)) for x in v) + ")"


def _table(path, tab):
    with open(path, "w") as f:
        f.write("(\n")
        for t, v in zip(tab.t, tab.v):
            f.write(f"    ( {t!r} {_vec(v)} )\n")
        f.write(")\n")


def _patch(bc, caseDir, tag):
    """One boundaryField entry, in the reference's own keywords."""
    def series(key, tab):
        name = f"timeVs_{tag}_{key}"
        _table(os.path.join(caseDir, "constant", name), tab)
        return f'        {key}\n        {{\n            "fileName|file" "$FOAM_CASE/constant/{name}";\n            outOfBounds clamp;\n        }}\n'
    s = f"        type {type(bc).__name__};\n"
    if isinstance(bc, PF.fixedValue):
        s += f"        value uniform {_vec(bc.value)};\n"
    elif isinstance(bc, PF.fixedDisplacement):
        s += f"        value uniform {_vec(bc.value)};\n" + (series("displacementSeries", bc.displacementSeries) if bc.displacementSeries else "")
    elif isinstance(bc, PF.fixedRotation):
        s += f"        value uniform {_vec(bc.value)};\n" + (series("thetaSeries", bc.thetaSeries) if bc.thetaSeries else "")
    elif isinstance(bc, PF.forceBeamDisplacementNR):
        s += f"        force uniform {_vec(bc.force)};\n" + (series("forceSeries", bc.forceSeries) if bc.forceSeries else "")
    elif isinstance(bc, PF.followerForceBeamDisplacementNR):
        s += series("followerForceSeries", bc.followerForceSeries)
    elif isinstance(bc, PF.linearSpringForceBeamDisplacementNR):
        s += (f"        linearSpringCoeffs\n        {{\n            springStiffness {bc.springStiffness!r};\n"
              f"            springDirection {_vec(bc.springDirection)};\n        }}\n")
        s += series("offsetForceSeries", bc.offsetForceSeries) if bc.offsetForceSeries else ""
    elif isinstance(bc, PF.momentBeamRotationNR):
        s += f"        moment uniform {_vec(bc.moment)};\n" + (series("momentSeries", bc.momentSeries) if bc.momentSeries else "")
    else:
        raise TypeError(type(bc))
    return s


def write_case(case, caseDir, rotateAngle=None):
    """BeamCase -> OpenFOAM case directory (circle/rectangle sections are passed as rectangle b, h or circle radius
    through ``case.section``, a (model, dict) pair attached by the test)."""
    for d in ("0", "constant", "system"):
        os.makedirs(os.path.join(caseDir, d), exist_ok=True)
    B = case.nBeams
    model, sec = case.section
    with open(os.path.join(caseDir, "constant", "beamProperties"), "w") as f:
        f.write(HEADER.format(cls="dictionary", obj="beamProperties"))
        f.write("beamModel coupledTotalLagNewtonRaphsonBeam;\nbeams\n(\n")
        for b in range(B):
            f.write(f"    beam_{b}\n    {{\n        crossSectionModel {model};\n        {model}CrossSectionModelDict\n        {{\n")
            for k, v in sec.items():
                f.write(f"            {k} {v!r};\n")
            f.write("        }\n")
            f.write(f"        length {float(case.per_beam(case.length)[b])!r};\n        nSegments {int(case.per_beam(case.nSegments, np.int32)[b])};\n")
            f.write(f"        E {float(case.per_beam(case.E)[b])!r};\n        G {float(case.per_beam(case.G)[b])!r};\n        rho {float(case.per_beam(case.rho)[b])!r};\n    }}\n")
        f.write(");\n\"coupledTotalLagNewtonRaphsonBeamCoeffs\"\n{\n")
        f.write(f"    nCorrectors {case.nCorrectors};\n    solutionTol {case.solutionTol!r};\n    absoluteTol {case.absoluteTol!r};\n    divergenceTol {case.divergenceTol!r};\n")
        if case.residualTol >= 0:
            f.write(f"    residualTol {case.residualTol!r};\n")
        if case.convergenceTol is not None:
            f.write(f"    convergenceTol {case.convergenceTol!r};\n")
        if case.scalingInertiaTensor != 1.0:
            f.write(f"    scalingInertiaTensor {case.scalingInertiaTensor!r};\n")
        f.write("    debug no;\n}\n")
    with open(os.path.join(caseDir, "constant", "g"), "w") as f:
        f.write(HEADER.format(cls="uniformDimensionedVectorField", obj="g"))
        f.write(f"dimensions      [0 1 -2 0 0 0 0];\nvalue           {_vec(case.g)};\n")
    for var, left, right in (("W", case.W_left, case.W_right), ("Theta", case.Theta_left, case.Theta_right)):
        with open(os.path.join(caseDir, "0", var), "w") as f:
            f.write(HEADER.format(cls="volVectorField", obj=var))
            f.write("dimensions      [0 1 0 0 0 0 0];\ninternalField   uniform (0 0 0);\nboundaryField\n{\n")
            for b in range(B):
                for side, bc in (("left", left), ("right", right)):
                    name = side if B == 1 else f"{side}_{b}"
                    f.write(f"    // beam {b}\n    {name}\n    {{\n{_patch(bc, caseDir, f'{var}_{name}')}    }}\n")
            f.write("    beam\n    {\n        type empty;\n    }\n}\n")
    if case.q is not None:
        with open(os.path.join(caseDir, "0", "q"), "w") as f:
            f.write(HEADER.format(cls="volVectorField", obj="q"))
            f.write(f"dimensions      [1 0 -2 0 0 0 0];\ninternalField   uniform {_vec(case.q)};\n"
                    "boundaryField\n{\n    \"right.*|left.*\"\n    {\n        type calculated;\n        value uniform (0 0 0);\n    }\n}\n")
    with open(os.path.join(caseDir, "system", "controlDict"), "w") as f:
        f.write(HEADER.format(cls="dictionary", obj="controlDict"))
        f.write(f"application     beamFoam;\nstartTime       0;\nendTime         {case.endTime!r};\ndeltaT          {case.deltaT!r};//comment\n"
                "writeControl    timeStep;\nwriteInterval   3;\nwritePrecision  10;\n"
                "functions\n{\n   beamDisplacements1\n   {\n       type    beamDisplacements;\n       historyPatch    right;\n   }\n}\n")
    with open(os.path.join(caseDir, "system", "fvSchemes"), "w") as f:
        f.write(HEADER.format(cls="dictionary", obj="fvSchemes"))
        f.write(f"ddtSchemes\n{{\n    default {case.d2dt2Scheme};\n}}\nd2dt2Schemes\n{{\n    default      {case.d2dt2Scheme};\n}}\n"
                "laplacianSchemes\n{\n    default         none;}\nsnGradSchemes\n{\n    default         orthogonal;\n}\n")
    if case.momentumContributions:
        with open(os.path.join(caseDir, "constant", "beamMomentumContributionProperties"), "w") as f:
            f.write(HEADER.format(cls="dictionary", obj="beamMomentumContributionProperties"))
            f.write("beamMomentumContributions\n(\n")
            for k, mc in enumerate(case.momentumContributions):
                t = type(mc).__name__.replace("Contribution", "")
                f.write(f"    contribution{k}\n    {{\n        beamMomentumContributionType {t};\n        {t}Coeffs\n        {{\n")
                for key, val in dataclasses.asdict(mc).items():
                    f.write(f"            {key} {val!r};\n")
                f.write("        }\n    }\n")
            f.write(");\n")
    with open(os.path.join(caseDir, "Allrun"), "w") as f:
        f.write("#!/bin/bash\n. $WM_PROJECT_DIR/bin/tools/RunFunctions\napplication=`getApplication`\n\nrunApplication createBeamMesh\n")
        if rotateAngle is not None:
            axis, deg = rotateAngle
            for b in range(B):
                tr = np.broadcast_to(np.zeros(3) if case.initialTranslation is None else case.initialTranslation, (B, 3))[b]
                f.write(f"runApplication -s beam_{b} setInitialPositionBeam -cellZone beam_{b} -translate '({float(tr[0])!r} {float(tr[1])!r} {float(tr[2])!r})' "
                        f"-rotateAngle '(({axis[0]} {axis[1]} {axis[2]}) {deg})'\n")
        f.write("\n# run beamFoam\nrunApplication $application\n")


def run_host(exe, caseDir, lib, steps):
    r = subprocess.run([exe, "Allrun", "-case", caseDir, "-lib", lib, "-steps", str(steps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    conv = np.loadtxt(os.path.join(caseDir, "history", "beamConvergenceData.dat"), ndmin=2)
    disp = np.loadtxt(os.path.join(caseDir, "history", "beamDisplacements_right.dat"), ndmin=2)
    return conv, disp


def read_written_field(caseDir, steps, deltaT, name):
    """internalField of <case>/<time>/<name> at the last write time (writeInterval 3), or None if none was due."""
    k = (steps // 3) * 3
    if k == 0:
        return None
    path = os.path.join(caseDir, f"{k * deltaT:.6g}", name)
    assert os.path.exists(path), path
    vals, inside = [], False
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("internalField"):
            inside = True
            continue
        if inside and ln.startswith("(") and ln.endswith(")") and len(ln) > 2:
            vals.append([float(x) for x in ln[1:-1].split()])
        elif inside and ln == ";":
            break
    return k, np.array(vals)


def run_python(case, lib, steps):
    m = coupledTotalLagNewtonRaphsonBeam(case, library=lib)
    rows, its = [], []
    def cb(mm):
        w, t = mm.tip(1)
        rows.append(np.concatenate([w, t], axis=1).reshape(-1))
        its.append((mm.iOuterCorr(), mm.lastStep.initialResidualNorm))
    m.run(nSteps=steps, callback=cb)
    m.close()
    return np.array(its), np.array(rows)


def _cases():
    c1 = cases.multiBeamDensity(3, 12)
    c1.section = ("rectangle", dict(b=0.2, h=0.2, nx=1, ny=1))
    c2 = cases.cantileverSpring()
    c2.section = ("circle", dict(radius=0.2))
    c3 = cases.cantileverFollowerForce(nSegments=16, deltaT=0.01)
    c3.section = ("circle", dict(radius=0.65828153))
    c4 = cases.cantilever(nSegments=20, deltaT=0.02)
    c4.section = ("circle", dict(radius=0.2))
    from beamfoam_b200.beamMomentumContribution import groundContactContribution, morisonDragContribution
    c5 = dataclasses.replace(cases.multiBeamDensity(2, 10), R=0.1, deltaT=2e-3,
                             initialTranslation=np.array([[0.0, 0.0, 0.1], [0.0, 0.4, 0.1]]),
                             momentumContributions=[morisonDragContribution(1.2, 0.05, 1000.0),
                                                    groundContactContribution(2e3, 50.0, 4e2, 0.3, 1.2)])
    c5.section = ("rectangle", dict(b=0.2, h=0.2, nx=1, ny=1))
    c5.name = "withMomentumContributions"
    return [(c1, ((0, -1, 0), 90), 6), (c2, None, 1), (c3, None, 5), (c4, None, 5), (c5, ((0, -1, 0), 90), 6)]


def _compare(exe, libPath, lib, tmp_path, tol):
    for case, rot, steps in _cases():
        d = str(tmp_path / case.name)
        write_case(case, d, rot)
        conv, disp = run_host(exe, d, libPath, steps)
        its, rows = run_python(case, lib, steps)
        assert conv.shape[0] == steps and np.array_equal(conv[:, 1].astype(int), its[:, 0].astype(int)), case.name
        scale = np.max(np.abs(rows))
        assert np.max(np.abs(disp[:, 1:] - rows)) <= tol * scale, case.name
        np.testing.assert_allclose(conv[:, 2], its[:, 1], rtol=1e-9)
        written = read_written_field(d, steps, case.deltaT, "W")  # runTime.write(): <time>/W in OpenFOAM format
        if written is not None:
            k, Wf = written
            m = coupledTotalLagNewtonRaphsonBeam(case, library=lib)
            m.run(nSteps=k)
            Wp = m.get_field("W")[0]
            m.close()
            assert Wf.shape == Wp.shape and np.max(np.abs(Wf - Wp)) <= 1e-8 * max(np.max(np.abs(Wp)), 1e-300), case.name


def test_cpp_host_matches_python_host_on_written_cases(host_exe, oracle_lib_path, tmp_path):
    _compare(host_exe, oracle_lib_path, load_library(oracle_lib_path), tmp_path, 1e-12)


def test_cpp_host_reports_reference_errors(host_exe, oracle_lib_path, tmp_path):
    case, rot, _ = _cases()[0]
    d = str(tmp_path / "bad")
    write_case(case, d, rot)
    # unknown d2dt2Scheme: CTL.C:708-720
    p = os.path.join(d, "system", "fvSchemes")
    text = open(p).read().replace("Newmark", "CrankNicolson")
    open(p, "w").write(text)
    r = subprocess.run([host_exe, "beamFoam", "-case", d, "-lib", oracle_lib_path], capture_output=True, text=True)
    assert r.returncode != 0 and "Please specify the d2dt2Schemes" in r.stderr
    open(p, "w").write(text.replace("CrankNicolson", "Newmark"))
    # a missing solver library is fatal: there is no CPU fallback
    r = subprocess.run([host_exe, "beamFoam", "-case", d, "-lib", str(tmp_path / "nope.so")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.skipif(not os.path.isdir(REF_TUTORIALS), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name,steps", [("cantilever", 5), ("cantileverFollowerForce", 5), ("largeDeflectionCantilever", 1),
                                        ("3dDynamicCantilever", 5), ("3DdynamicCantileverMultiBeamDensity", 5),
                                        ("freeFlexibleBeam", 5), ("beamAndSpringInSeries", 3), ("cantileverSpring", 1)])
def test_cpp_host_runs_shipped_tutorials_unchanged(host_exe, oracle_lib_path, tmp_path, name, steps):
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(REF_TUTORIALS, name), d)
    conv, disp = run_host(host_exe, d, oracle_lib_path, steps)
    case = cases.TUTORIALS[name]()
    its, rows = run_python(case, load_library(oracle_lib_path), steps)
    assert np.array_equal(conv[:, 1].astype(int), its[:, 0].astype(int))
    assert np.max(np.abs(disp[:, 1:] - rows)) <= 1e-9 * max(np.max(np.abs(rows)), 1e-300)


@pytest.mark.gpu
def test_cpp_host_on_gpu_library(host_exe, tmp_path):
    from beamfoam_b200 import build
    _compare(host_exe, build.OUTPUT, load_library(), tmp_path, 1e-12)
