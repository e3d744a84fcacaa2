"""Case descriptions: replicas of the reference tutorials and the synthetic bundles
of BASELINE.json's configs.

Each function restates the dictionaries of one tutorial case
(tutorials/<name>/{constant/beamProperties, 0/W, 0/Theta, system/fvSchemes,
system/controlDict, constant/timeVs*, Allrun}); SURVEY.md Appendix A is the map.
"""
from __future__ import annotations

import numpy as np

from .beamModel import BeamCase, axisAngleRotation
from .crossSections import circle, rectangle
from .fvPatchFields import (fixedDisplacement, fixedRotation, fixedValue, followerForceBeamDisplacementNR,
                            forceBeamDisplacementNR, interpolationTable, linearSpringForceBeamDisplacementNR,
                            momentBeamRotationNR)


def _section(cs):
    # beamModel.H:449-489: Iyy = crossSection.Ixx(), Izz = crossSection.Iyy(), J = IT()
    return dict(A=cs.A(), Iyy=cs.Ixx(), Izz=cs.Iyy(), J=cs.IT())


def cantilever(nSegments: int = 50, deltaT: float = 0.01) -> BeamCase:
    """tutorials/cantilever: roll-up by a prescribed end rotation ramp to (0 -2pi 0);
    zero tip force; 50 segments; 100 steady load steps; convergenceTol 1e-10."""
    return BeamCase(
        name="cantilever", nSegments=nSegments, length=10.0, **_section(circle(0.2)),
        E=7.9577472e4, G=3.9788736e4,
        W_left=fixedValue(variable="W"),
        W_right=forceBeamDisplacementNR(forceSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 0))])),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=fixedRotation(thetaSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, -6.28318, 0))])),
        d2dt2Scheme="steadyState", nCorrectors=200, convergenceTol=1e-10, deltaT=deltaT, endTime=1.0)


def cantileverFollowerForce(nSegments: int = 40, deltaT: float = 0.001) -> BeamCase:
    """tutorials/cantileverFollowerForce: follower tip force ramp to (0 0 1.34e5) in the
    material frame, zero tip moment; residualTol 1e-8, solutionTol 1e-10; 1000 steps.
    ``nSegments=10000`` is BASELINE.json configs[1]."""
    return BeamCase(
        name="cantileverFollowerForce", nSegments=nSegments, length=100.0, **_section(circle(0.65828153)),
        E=2.3731844e8, G=1.1865922e8,
        W_left=fixedValue(variable="W"),
        W_right=followerForceBeamDisplacementNR(
            followerForceSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 1.34e5))])),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=momentBeamRotationNR(momentSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 0))])),
        d2dt2Scheme="steadyState", nCorrectors=1000, residualTol=1e-8, solutionTol=1e-10,
        deltaT=deltaT, endTime=1.0)


def largeDeflectionCantilever(tipForce: float = 0.588, nSegments: int = 80) -> BeamCase:
    """tutorials/largeDeflectionCantilever: self-weight (rho 7730, g (0 0 -9.81)) plus a tip
    force (0 0 -tipForce) applied in ONE steady step; thin steel strip 0.025 x 0.0004."""
    return BeamCase(
        name="largeDeflectionCantilever", nSegments=nSegments, length=0.4,
        **_section(rectangle(b=0.025, h=0.0004)), E=194.3e9, G=74.73e9, rho=7730.0,
        W_left=fixedValue(variable="W"),
        W_right=forceBeamDisplacementNR(
            forceSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, -tipForce))])),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=momentBeamRotationNR(moment=(0, 0, 0)),
        g=(0.0, 0.0, -9.81), d2dt2Scheme="steadyState", nCorrectors=200, residualTol=1e-8,
        solutionTol=1e-8, absoluteTol=1e-8, convergenceTol=1e-8, divergenceTol=1e8,
        deltaT=1.0, endTime=1.0)


def dynamicCantilever(nBeams: int = 1, nSegments: int = 20, rho=1000.0, deltaT: float = 1e-3,
                      loadScale=None, name: str = "3dDynamicCantilever") -> BeamCase:
    """tutorials/3dDynamicCantilever (nBeams=1, n=20, rho 1000) and
    tutorials/3DdynamicCantileverMultiBeamDensity (3 beams, n=40, rho 500/1000/2000):
    Newmark (beta .25, gamma .5), square 0.2 x 0.2, L = 2, initial configuration rotated
    90 deg about (0 -1 0), left end clamped, right end constant force (2000 2000 0) and
    zero moment, nCorrectors 200, solutionTol 1e-10 (=> residualTol 1e-8)."""
    return BeamCase(
        name=name, nBeams=nBeams, nSegments=nSegments, length=2.0, **_section(rectangle(b=0.2, h=0.2)),
        E=15.293e6, G=5.882e6, rho=rho,
        W_left=fixedValue(variable="W"),
        W_right=forceBeamDisplacementNR(
            forceSeries=interpolationTable([(0, (2000, 2000, 0)), (1, (2000, 2000, 0))]), scale=loadScale),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=momentBeamRotationNR(momentSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 0))])),
        initialRotation=axisAngleRotation((0, -1, 0), 90.0),
        d2dt2Scheme="Newmark", nCorrectors=200, solutionTol=1e-10, deltaT=deltaT, endTime=1.0)


def multiBeamDensity(nBeams: int = 3, nSegments: int = 40, jitter: bool = False) -> BeamCase:
    """3DdynamicCantileverMultiBeamDensity scaled to ``nBeams`` beams (BASELINE.json
    configs[2]/[3]: 4096 and 65,536 beams x 200 cells): rho cycles 500/1000/2000 by
    ``b mod 3``.  The per-beam translation (0 0.4b 0) of the Allrun only moves refWf,
    which the hot path never reads.  ``jitter`` scales each beam's load by 1 + 0.1*u_b,
    u_b = default_rng(20261019).random(nBeams)."""
    rho = np.array([500.0, 1000.0, 2000.0])[np.arange(nBeams) % 3]
    scale = None
    if jitter:
        scale = 1.0 + 0.1 * np.random.default_rng(20261019).random(nBeams)
    return dynamicCantilever(nBeams=nBeams, nSegments=nSegments, rho=rho, loadScale=scale,
                             name=f"multiBeamDensity_{nBeams}x{nSegments}")


def freeFlexibleBeam() -> BeamCase:
    """tutorials/freeFlexibleBeam: free-free beam, Newmark, dt 0.01, 4000 steps,
    scalingInertiaTensor 200.39427, rotated 53.130102 deg about (0 0 -1)."""
    z = (0, 0, 0)
    return BeamCase(
        name="freeFlexibleBeam", nSegments=10, length=10.0, **_section(circle(0.44721360)),
        E=1.5915494e4, G=1.5915494e4, rho=1.5915466,
        W_left=forceBeamDisplacementNR(forceSeries=interpolationTable([(0, z), (1000, z)])),
        W_right=forceBeamDisplacementNR(forceSeries=interpolationTable(
            [(0, (8, 0, 0)), (0.1, (8, 0, 0)), (2.5, (8, 0, 0)), (2.51, z), (2.6, z), (1000, z)])),
        Theta_left=momentBeamRotationNR(momentSeries=interpolationTable([(0, z), (1000, z)])),
        Theta_right=momentBeamRotationNR(momentSeries=interpolationTable(
            [(0, (56.568542495, 56.568542495, 0)), (0.1, (56.568542495, 56.568542495, 0)),
             (2.5, (56.568542495, 56.568542495, 0)), (2.51, z), (2.6, z), (1000, z)])),
        initialRotation=axisAngleRotation((0, 0, -1), 53.130102),
        d2dt2Scheme="Newmark", nCorrectors=200, solutionTol=1e-10, scalingInertiaTensor=200.39427,
        deltaT=0.01, endTime=40.0)


def beamAndSpringInSeries() -> BeamCase:
    """tutorials/beamAndSpringInSeries: left end pulled by displacementSeries to (-1 0 0),
    right end on a linear spring k=1000 along x with offset force (100 0 0)."""
    return BeamCase(
        name="beamAndSpringInSeries", nSegments=50, length=10.0, **_section(circle(0.2)),
        E=2e11, G=1e11,
        W_left=fixedDisplacement(displacementSeries=interpolationTable([(0, (0, 0, 0)), (1, (-1, 0, 0))])),
        W_right=linearSpringForceBeamDisplacementNR(
            springStiffness=1000.0, springDirection=(1, 0, 0),
            offsetForceSeries=interpolationTable([(0, (100, 0, 0)), (1, (100, 0, 0))])),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=momentBeamRotationNR(momentSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 0))])),
        d2dt2Scheme="steadyState", nCorrectors=2000, residualTol=1e-8, solutionTol=1e-10,
        deltaT=0.1, endTime=1.0)


def cantileverSpring() -> BeamCase:
    """tutorials/cantileverSpring: clamped beam under q = (0 -1000 0) N/m resting on a
    spring k=800 along y at the right end; one steady step."""
    return BeamCase(
        name="cantileverSpring", nSegments=50, length=10.0, **_section(circle(0.2)),
        E=2e8, G=1e8,
        W_left=fixedValue(variable="W"),
        W_right=linearSpringForceBeamDisplacementNR(springStiffness=800.0, springDirection=(0, 1, 0)),
        Theta_left=fixedValue(variable="Theta"),
        Theta_right=momentBeamRotationNR(momentSeries=interpolationTable([(0, (0, 0, 0)), (1, (0, 0, 0))])),
        q=(0.0, -1000.0, 0.0),
        d2dt2Scheme="steadyState", nCorrectors=2000, residualTol=1e-8, solutionTol=1e-10,
        deltaT=1.0, endTime=1.0)


TUTORIALS = {
    "cantilever": cantilever,
    "cantileverFollowerForce": cantileverFollowerForce,
    "largeDeflectionCantilever": largeDeflectionCantilever,
    "3dDynamicCantilever": dynamicCantilever,
    "3DdynamicCantileverMultiBeamDensity": multiBeamDensity,
    "freeFlexibleBeam": freeFlexibleBeam,
    "beamAndSpringInSeries": beamAndSpringInSeries,
    "cantileverSpring": cantileverSpring,
}
