"""Cross-section constants consumed once at construction (host side).

Only the six scalars the hot path needs are restated (A, Ixx, Iyy, IT); the
reference's crossSectionModel hierarchy is OUT OF SCOPE.
Formulas: src/wireBunchingModels/beamModels/beamModel/crossSections/circle/circle.H:123-144
and .../rectangle/rectangle.H:145-166.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass(frozen=True)
class circle:
    radius: float
    scalingArea: float = 1.0
    scalingSecondMomentArea: float = 1.0

    def R(self):
        return self.radius

    def A(self):
        return self.scalingArea * math.pi * self.radius ** 2

    def Ixx(self):
        return self.scalingSecondMomentArea * math.pi * self.radius ** 4 / 4

    def Iyy(self):
        return self.scalingSecondMomentArea * math.pi * self.radius ** 4 / 4

    def IT(self):
        return self.scalingSecondMomentArea * math.pi * self.radius ** 4 / 2


@dataclass(frozen=True)
class rectangle:
    b: float
    h: float
    scalingArea: float = 1.0
    scalingSecondMomentArea: float = 1.0

    def R(self):  # rectangle.H:133-136
        return self.h / 2

    def A(self):
        return self.scalingArea * self.b * self.h

    def Ixx(self):
        return self.scalingSecondMomentArea * self.b * self.h ** 3 / 12

    def Iyy(self):
        return self.scalingSecondMomentArea * self.h * self.b ** 3 / 12

    def IT(self):
        return self.Ixx() + self.Iyy()
