"""Host-side mirrors of the reference's run-time selectable ``beamMomentumContribution`` classes
(src/wireBunchingModels/beamMomentumContribution; listed in constant/beamMomentumContributionProperties,
CTL.C:1137-1182).  Only the dictionary data lives here; the per-cell arithmetic (Hermite-spline mid-point
derivative, drag / penalty contact forces; Evolve.C:539-571) runs behind the C ABI
(``bf_set_momentum_contributions``)."""
from __future__ import annotations

from dataclasses import dataclass

from . import _lib as L


@dataclass(frozen=True)
class morisonDragContribution:
    """morisonDragContribution.C:48-75 (``<name>Coeffs { Cdn; Cdt; rhoFluid; }``)."""
    Cdn: float
    Cdt: float
    rhoFluid: float
    type = L.BF_CONTRIB_MORISON_DRAG

    def coeffs(self):
        return (self.Cdn, self.Cdt, self.rhoFluid)


@dataclass(frozen=True)
class groundContactContribution:
    """groundContactContribution.C:48-83 (``<name>Coeffs { kNormal; cNormal; kTangential; muFriction; groundZ; }``)."""
    kNormal: float
    cNormal: float
    kTangential: float
    muFriction: float
    groundZ: float
    type = L.BF_CONTRIB_GROUND_CONTACT

    def coeffs(self):
        return (self.kNormal, self.cNormal, self.kTangential, self.muFriction, self.groundZ)
