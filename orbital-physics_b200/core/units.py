"""Scalar unit wrappers used by the Kepler-element model (host only).

API-compatible with the reference's core/units.py:1-86: a `Unit` carries
`.value` (float) and `.unit` (tag); angles normalise modulo a full turn on
construction; `+`/`-` re-wrap through the subclass so angle arithmetic
re-normalises.  Conversions come from one table instead of hand-written methods.
"""
from __future__ import annotations

import math

AU_METERS = 1.495978707e11
KG_SOLAR = 1.98847e30
_SECONDS_PER_DAY = 86400.0


class Unit:
    def __init__(self, value: "float | int", unit: str):
        self.value = float(value)
        self.unit = unit

    def __repr__(self):
        return f"{self.unit.upper()}({self.value})"

    def _combine(self, other, sign: float, verb: str):
        if self.unit != other.unit:
            raise ValueError(f"Cannot {verb} objects of different types.")
        return type(self)(self.value + sign * other.value)

    def __add__(self, other):
        return self._combine(other, 1.0, "add")

    def __sub__(self, other):
        return self._combine(other, -1.0, "subtract")


def _unit_type(cls_name: str, tag: str, wrap=None):
    """Build a one-argument Unit subclass; `wrap` optionally normalises the value."""

    def __init__(self, value):
        Unit.__init__(self, wrap(value) if wrap else value, tag)

    return type(cls_name, (Unit,), {"__init__": __init__, "__doc__": f"A quantity in {tag}."})


Radians = _unit_type("Radians", "radians", lambda v: v % (2 * math.pi))
Degrees = _unit_type("Degrees", "degrees", lambda v: v % 360)
Meters = _unit_type("Meters", "meters")
AU = _unit_type("AU", "au")
Kilograms = _unit_type("Kilograms", "kilograms")
SolarMasses = _unit_type("SolarMasses", "m_solar")
Seconds = _unit_type("Seconds", "seconds")
Days = _unit_type("Days", "days")


def _converter(target, fn):
    def convert(self):
        return target(fn(self.value))
    return convert


# (source type, method name, target type, value map)
for _src, _name, _dst, _fn in (
    (Radians, "to_degrees", Degrees, math.degrees),
    (Degrees, "to_radians", Radians, math.radians),
    (Meters, "to_au", AU, lambda v: v / AU_METERS),
    (AU, "to_meters", Meters, lambda v: v * AU_METERS),
    (Kilograms, "to_solar_masses", SolarMasses, lambda v: v / KG_SOLAR),
    (SolarMasses, "to_kilograms", Kilograms, lambda v: v * KG_SOLAR),
    (Seconds, "to_days", Days, lambda v: v / _SECONDS_PER_DAY),
    (Days, "to_seconds", Seconds, lambda v: v * _SECONDS_PER_DAY),
):
    setattr(_src, _name, _converter(_dst, _fn))
del _src, _name, _dst, _fn
