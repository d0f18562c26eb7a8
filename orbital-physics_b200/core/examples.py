"""Canned scenarios (two-body, Sun-Earth-Moon, Lagrange triangle, solar system from Kepler elements).

Same entry points, defaults and physical set-up as the reference's
core/examples.py:11-233; they double as parity scenarios.  Plotting / video
export is skipped (with a note) when matplotlib is unavailable, instead of
failing at import time.  As in the reference, `unit_profile` is resolved but the
engine integrates in SI regardless.
"""
from __future__ import annotations

import numpy as np

from core.constants import UnitSystem, get_unit_profile
from core.engine import SimulationEngine, run_simulation
from core.physics import Coordinates, Object, ObjectCollection, set_circular_orbit
from core.plot import plot_orbits, render_orbital_mp4

EARTH = dict(mass=5.972e24, radius=6.371e6)
MOON = dict(mass=7.348e22, radius=1.737e6)
SUN = dict(mass=1.98847e30, radius=6.9634e8)
EARTH_MOON_DISTANCE = 384400e3
ASTRONOMICAL_UNIT = 1.495978707e11


def _try(render, *args, **kwargs):
    try:
        return render(*args, **kwargs)
    except RuntimeError as exc:
        print(f"[examples] skipped rendering: {exc}")


def two_body_problem(body1_mass: float = EARTH["mass"], body1_radius: float = EARTH["radius"],
                     body2_mass: float = MOON["mass"], body2_radius: float = MOON["radius"],
                     distance: float = EARTH_MOON_DISTANCE, dt: float = 60 * 60, steps: int = 1000,
                     unit_profile: UnitSystem = "si"):
    """Two bodies on a circular orbit about their barycentre (zero net momentum)."""
    unit_profile = get_unit_profile(unit_profile)
    primary = Object(mass=body1_mass, radius=body1_radius, velocity=np.zeros(3), coordinates=Coordinates(0, 0, 0))
    secondary = Object(mass=body2_mass, radius=body2_radius, velocity=np.zeros(3),
                       coordinates=Coordinates(distance, 0, 0))
    set_circular_orbit(primary=primary, secondary=secondary)
    collection = ObjectCollection([primary, secondary])
    for obj in collection:
        print(obj)
    engine = SimulationEngine(collection, dt=dt, softening=1e3, restitution=1.0)
    run_simulation(engine, steps=steps)
    _try(plot_orbits, engine, every_n=5, plane="xy", separate=False, with_velocity=False)
    return engine


def sun_earth_moon(steps: int = 5000, dt: float = 3600., moon_incl_deg: float = 0.0, softening: float = 1e3,
                   unit_profile: UnitSystem = "si"):
    """Earth-Moon pair orbiting the Sun; all velocities are assigned fp64 arrays."""
    unit_profile = get_unit_profile(unit_profile)
    AU, R_em = ASTRONOMICAL_UNIT, EARTH_MOON_DISTANCE
    sun = Object(SUN["mass"], SUN["radius"], velocity=np.zeros(3), coordinates=Coordinates(0, 0, 0))
    earth = Object(EARTH["mass"], EARTH["radius"], velocity=np.zeros(3), coordinates=Coordinates(AU, 0, 0))
    moon_pos = np.array([AU + R_em, 0.0, 0.0])
    if abs(moon_incl_deg) > 0:
        moon_pos = np.array([AU + R_em, 0.0, R_em * np.sin(np.deg2rad(moon_incl_deg))])
    moon = Object(MOON["mass"], MOON["radius"], velocity=np.zeros(3), coordinates=Coordinates.from_iterable(moon_pos))

    set_circular_orbit(sun, earth)                  # Sun + Earth: circular, zero net momentum
    v_cm = earth.velocity.copy()                    # heliocentric velocity wanted for the Earth-Moon barycentre

    sep = moon.position() - earth.position()
    R = np.linalg.norm(sep)
    r_hat = sep / R
    t_hat = np.cross(np.array([0.0, 0.0, 1.0]), r_hat)       # prograde tangent
    if np.linalg.norm(t_hat) < 1e-12:
        t_hat = np.array([0.0, 1.0, 0.0])
    t_hat /= np.linalg.norm(t_hat)
    m_e, m_m = EARTH["mass"], MOON["mass"]
    v_rel = np.sqrt(unit_profile.G * (m_e + m_m) / R) * t_hat
    earth.velocity = v_cm - (m_m / (m_e + m_m)) * v_rel
    moon.velocity = v_cm + (m_e / (m_e + m_m)) * v_rel

    engine = SimulationEngine(ObjectCollection([sun, earth, moon]), dt=dt, softening=softening, restitution=1.0)
    run_simulation(engine, steps=steps, print_every=500)
    _try(plot_orbits, engine, every_n=10, plane="xy", separate=False, with_velocity=False, show_barycenter=True,
         barycenter_trail=True)
    return engine


def three_body_equilateral(m: float = 1e22, R: float = 1e7, dt: float = 50.0, steps: int = 8000,
                           softening: float = 1e3, unit_profile: UnitSystem = "si",
                           out_path: str = "three_body_equilateral.mp4"):
    """Lagrange's rigidly rotating equilateral triangle of three equal masses (fp32 velocities via the ctor)."""
    unit_profile = get_unit_profile(unit_profile)
    half, h = -0.5 * R, np.sqrt(3) / 2 * R
    vertices = [np.array([R, 0.0, 0.0]), np.array([half, h, 0.0]), np.array([half, -h, 0.0])]
    z_hat = np.array([0.0, 0.0, 1.0])
    tangents = [np.cross(z_hat, p / np.linalg.norm(p)) for p in vertices]
    speed = np.sqrt(unit_profile.G * m / (np.sqrt(3.0) * R))
    bodies = [Object(mass=m, radius=(m / 5000.0) ** (1 / 3), velocity=speed * tangents[k],
                     coordinates=Coordinates.from_iterable(vertices[k])) for k in range(3)]
    engine = SimulationEngine(ObjectCollection(bodies), dt=dt, softening=softening, restitution=1.0)
    run_simulation(engine, steps=steps, print_every=500)
    _try(render_orbital_mp4, engine, out_path=out_path, plane="xy", fps=30, duration_s=30, with_velocity=False,
         show_barycenter=True, barycenter_trail=True, every_n=5)
    return engine


def solar_system_objects(moons: bool = False, parent_offset: bool = False):
    """Objects for the Kepler-element dataset, converted to SI exactly as the reference's example/app do."""
    from core.datasets import solar_system_v2
    system = solar_system_v2(moons=moons)
    system.standardize_units(mass_unit="kilograms", distance_unit="meters", angle_unit="radians",
                             time_unit="seconds")
    bodies = []
    for body in system:
        r, v = body.get_state()
        if parent_offset and body.parent is not None:
            pr, pv = body.parent.get_state()
            r = np.array(pr) + np.array(r)
            v = np.array(pv) + np.array(v)
        bodies.append(Object(mass=body.mass.value, radius=body.radius.value, velocity=np.array(v, dtype=np.float64),
                             coordinates=Coordinates(*r), name=body.name))
    return bodies, system


def sol_from_kepler_dataset(out_path: str = "sol_from_keplerian.mp4", days: int = 365, dt: float = None,
                            print_every: int = 100):
    """Sun + planets + dwarf planets from core.datasets, one step per day by default."""
    dt = 86400.0 if dt is None else dt
    bodies, _ = solar_system_objects(moons=False)
    engine = SimulationEngine(ObjectCollection(bodies), dt=dt, softening=1e6, restitution=1.0)
    run_simulation(engine, steps=days, print_every=print_every)
    _try(render_orbital_mp4, engine, out_path=out_path, plane="xy", fps=30, duration_s=30, with_velocity=False,
         show_barycenter=True, barycenter_trail=True, every_n=5)
    return engine
