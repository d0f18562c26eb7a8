"""Keplerian-element body model: generates initial conditions.

`Body.get_state` is the host path (bit-identical to the reference); `batch_states` / `System.get_states`
evaluate many bodies in one device launch (liborbital_b200 `orb_kepler_states`, csrc/kepler.cu: same operation
order, device sin/cos, so within ~1e-15 relative of the host path rather than bit-identical).

API-compatible with the reference's core/body.py:14-317 (`Body`, `System`).
This is the step *before* the hot path: it runs once per body at start-up and
feeds `core.physics.Object`.  The arithmetic order of `get_state` follows the
reference expression by expression (core/body.py:184-249) so the SI state is
bit-identical -- tests/test_model.py pins it against golden vectors.
"""
from __future__ import annotations

import math

from core.constants import STANDARD
from core.physics import moment_of_inertia, solve_kepler
from core.units import AU, Days, Degrees, Kilograms, Meters, Radians, Seconds, SolarMasses, Unit

G = STANDARD.G


def _metres(q) -> float:
    return (q.to_meters() if isinstance(q, AU) else q).value


def _kilograms(q) -> float:
    return (q.to_kilograms() if isinstance(q, SolarMasses) else q).value


def _radians(q) -> float:
    return (q.to_radians() if isinstance(q, Degrees) else q).value


class Body:
    """Orbital elements of one body about its parent (planets: e, a, I, Omega, varpi, L;
    moons: e, a, I, Omega, omega, M), plus mass and radius."""

    _FIELDS = ("name", "a", "e", "I", "L", "long_peri", "long_node", "M", "arg_peri", "mass", "radius",
               "b", "mu", "fg", "T")

    def __init__(self, name, a, e, I, L, M, long_peri, long_node, arg_peri, mass, radius,
                 b=None, fg=None, T=None, mu=None, parent: "Body | None" = None):
        self.name = name
        self.a, self.e, self.I = a, e, I
        self.L, self.M = L, M
        self.long_peri, self.long_node, self.arg_peri = long_peri, long_node, arg_peri
        self.mass, self.radius = mass, radius
        self.b, self.fg = b, fg
        self.T = Seconds(T) if isinstance(T, float) else T
        self.parent = parent
        self.mu = mu
        self.derive()

    # -- derived quantities --------------------------------------------------
    def derive(self):
        """Fill in whichever of mu, b, (varpi | omega), (M | L), fg, T were not supplied."""
        if self.mu is None:
            self.mu = self.get_mu()
        if self.b is None:
            self.b = self.get_b()
        if self.long_peri is None:                      # varpi = Omega + omega
            assert self.arg_peri is not None, "Must provide either long_peri or arg_peri"
            self.long_peri = self.long_node + self.arg_peri
        elif self.arg_peri is None:
            self.arg_peri = self.long_peri - self.long_node
        if self.M is None:                              # L = varpi + M
            assert self.L is not None
            self.M = self.L - self.long_peri
        elif self.L is None:
            self.L = self.long_peri + self.M
        if self.fg is None:
            self.fg = self.get_fg()
        if self.T is None:
            self.T = self.get_T()

    def get_mu(self):
        """Standard gravitational parameter G*m."""
        return G * _kilograms(self.mass)

    def get_fg(self):
        """Surface gravity mu / r^2."""
        return self.mu / (_metres(self.radius) ** 2)

    def get_T(self):
        """Orbital period 2 pi sqrt(a^3 / (G M_parent)); None for a root body."""
        if self.parent is None:
            return None
        return Seconds(2 * math.pi * math.sqrt((_metres(self.a) ** 3) / (G * _kilograms(self.parent.mass))))

    def get_b(self):
        """Semi-minor axis a sqrt(1 - e^2)."""
        return Meters(_metres(self.a) * math.sqrt(1 - self.e ** 2))

    def mean_motion(self):
        """n = sqrt(mu_parent / a^3); 0 for a root body."""
        if self.parent is None:
            return 0.0
        return math.sqrt(self.parent.mu / _metres(self.a) ** 3)

    def rotational_intertia(self):
        """Moment of inertia of a uniform sphere with this mass and radius."""
        return moment_of_inertia(_kilograms(self.mass), _metres(self.radius), shape="sphere")

    # -- serialisation -------------------------------------------------------
    def to_dict(self):
        d = {k: getattr(self, k) for k in self._FIELDS}
        d["parent"] = self.parent.name if self.parent else ""
        return d

    def to_json(self) -> dict:
        return {k: (v.value if isinstance(v, Unit) else v) for k, v in self.to_dict().items()}

    def __repr__(self):
        return f"Body({self.to_dict()})"

    # -- elements -> Cartesian state ------------------------------------------
    def get_state(self):
        """Parent-relative position [m] and velocity [m/s] in the inertial frame."""
        if self.parent is None:
            return [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        M = _radians(self.M)
        a = _metres(self.a)
        inc = _radians(self.I)
        Omega = _radians(self.long_node)
        omega = _radians(self.arg_peri)
        b = _metres(self.b)
        n = self.mean_motion()
        e = self.e

        E = solve_kepler(M, e)
        cE, sE = math.cos(E), math.sin(E)
        # perifocal frame (z = 0)
        denom = 1 - e * cE
        r_pf = (a * (cE - e), b * sE, 0.0)
        v_pf = (-a * n * sE / denom, a * n * math.sqrt(1 - e ** 2) * cE / denom, 0.0)

        # R = Rz(Omega) Rx(i) Rz(omega)
        cw, sw = math.cos(omega), math.sin(omega)
        ci, si = math.cos(inc), math.sin(inc)
        cO, sO = math.cos(Omega), math.sin(Omega)
        R = (
            (cO * cw - sO * sw * ci, -cO * sw - sO * cw * ci, sO * si),
            (sO * cw + cO * sw * ci, -sO * sw + cO * cw * ci, -cO * si),
            (sw * si, cw * si, ci),
        )

        def rotate(p):
            return [row[0] * p[0] + row[1] * p[1] + row[2] * p[2] for row in R]

        return rotate(r_pf), rotate(v_pf)


def batch_states(bodies, device: int | None = None):
    """Parent-relative (r[n,3], v[n,3]) of many bodies in ONE device launch.

    Batched counterpart of a Python loop over `Body.get_state` (reference core/body.py:184-249); bodies without
    a parent sit at the origin, as there.  Raises core._native.NativeError without the CUDA library/device.
    """
    import numpy as np
    from core import _native
    from core.physics import default_device
    bodies = list(bodies)
    idx = [k for k, bd in enumerate(bodies) if bd.parent is not None]
    r, v = np.zeros((len(bodies), 3)), np.zeros((len(bodies), 3))
    if idx:
        sel = [bodies[k] for k in idx]
        cols = ([_radians(bd.M) for bd in sel], [bd.e for bd in sel], [_metres(bd.a) for bd in sel],
                [_metres(bd.b) for bd in sel], [bd.mean_motion() for bd in sel], [_radians(bd.I) for bd in sel],
                [_radians(bd.long_node) for bd in sel], [_radians(bd.arg_peri) for bd in sel])
        rr, vv = _native.kepler_states(*(np.array(c, dtype=np.float64) for c in cols),
                                       device=default_device() if device is None else device)
        r[idx], v[idx] = rr, vv
    return r, v


class System:
    """An ordered set of bodies plus the unit choice their elements are expressed in."""

    def get_states(self, device: int | None = None):
        """(r[n,3], v[n,3]) of every body, evaluated on the device in one launch (see `batch_states`)."""
        return batch_states(self.bodies, device)

    def __init__(self, bodies, distance_unit="meters", mass_unit="kg", angle_unit="radians", time_unit="seconds"):
        self.bodies = bodies
        self.distance_unit = distance_unit
        self.mass_unit = mass_unit
        self.angle_unit = angle_unit
        self.time_unit = time_unit

    def __getitem__(self, idx):
        return self.bodies[idx]

    def __len__(self):
        return len(self.bodies)

    def __repr__(self):
        return f"System({self.bodies})"

    def to_dict(self):
        return {b.name: b.to_dict() for b in self.bodies}

    def to_json(self):
        return {b.name: b.to_json() for b in self.bodies}

    def values(self):
        return self.to_json()

    def _convert(self, value):
        if not isinstance(value, Unit):
            return value
        rules = (
            (Meters, self.distance_unit, "au", "to_au"),
            (AU, self.distance_unit, "meters", "to_meters"),
            (Radians, self.angle_unit, "degrees", "to_degrees"),
            (Degrees, self.angle_unit, "radians", "to_radians"),
            (Kilograms, self.mass_unit, "m_solar", "to_solar_masses"),
            (SolarMasses, self.mass_unit, "kilograms", "to_kilograms"),
            (Seconds, self.time_unit, "days", "to_days"),
            (Days, self.time_unit, "seconds", "to_seconds"),
        )
        for kind, current, wanted, method in rules:
            if isinstance(value, kind) and current == wanted:
                return getattr(value, method)()
        return value

    def standardize_units(self, distance_unit=None, mass_unit=None, angle_unit=None, time_unit=None):
        """Convert, in place, every Unit-typed attribute of every body to the requested units."""
        self.distance_unit = distance_unit or self.distance_unit
        self.mass_unit = mass_unit or self.mass_unit
        self.angle_unit = angle_unit or self.angle_unit
        self.time_unit = time_unit or self.time_unit
        for body in self.bodies:
            for attr, val in list(body.__dict__.items()):
                setattr(body, attr, self._convert(val))
