"""Physical constants and unit profiles.

API-compatible with the reference's core/constants.py:1-80 (same names, same
values); only STANDARD.G is on the hot path (the engine always integrates in
SI, reference core/engine.py:41,78).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass

# metres per astronomical unit, seconds per day, J2000 epoch as a Julian date
AU = 1.495978707e11
DAY = 86400.0
JULIAN_DAY = 86400.0
J2000_JD = 2451545.0


class UnitSystem(str, enum.Enum):
    """Which base units a profile is expressed in."""

    ASTRO = "astro"   # AU, solar masses, days
    SI = "si"         # metres, kilograms, seconds


@dataclass(frozen=True)
class UnitProfile:
    """Gravitational constant plus the conversion anchors of one unit system."""

    name: UnitSystem
    G: float
    distance_unit: str
    mass_unit: str
    time_unit: str
    AU: float
    M_SUN: float
    DAY: float


_PROFILE_TABLE = {
    UnitSystem.ASTRO: dict(G=0.0002959122082855911, distance_unit="AU", mass_unit="M_sun", time_unit="day",
                           AU=1.0, M_SUN=1.0, DAY=1.0),
    UnitSystem.SI: dict(G=6.67430e-11, distance_unit="m", mass_unit="kg", time_unit="s",
                        AU=1.495978707e11, M_SUN=1.98847e30, DAY=86400.0),
}

ASTRO = UnitProfile(name=UnitSystem.ASTRO, **_PROFILE_TABLE[UnitSystem.ASTRO])
STANDARD = UnitProfile(name=UnitSystem.SI, **_PROFILE_TABLE[UnitSystem.SI])


@dataclass(frozen=True)
class IntegratorParams:
    """Default step / softening pair, in the units of the matching profile."""

    softening: float
    dt: float


DEFAULT_STANDARD_INTEGRATOR = IntegratorParams(dt=60 * 60, softening=1.0)
DEFAULT_ASTRO_INTEGRATOR = IntegratorParams(dt=1.0, softening=1e-6)


def get_unit_profile(name: "str | UnitSystem") -> UnitProfile:
    """Look a profile up by enum member or (case-insensitive) string."""
    key = UnitSystem(name.lower()) if isinstance(name, str) else name
    if key == UnitSystem.ASTRO:
        return ASTRO
    if key == UnitSystem.SI:
        return STANDARD
    raise ValueError(f"Unknown unit system: {name}")
