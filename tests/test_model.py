"""Host-side model layer (units, constants, Kepler elements -> state) vs the reference's outputs.

Golden: tests/golden/kepler.npz from core/body.py:65-97,184-249, core/datasets.py:13-56,
core/physics.py:43-71 of the reference.  Bar: bit-exact (these feed the initial condition).
"""
import math

import numpy as np
import pytest


def test_dataset_states_bit_exact(golden):
    from core.datasets import solar_system, solar_system_v2
    g = golden("kepler")
    assert solar_system is solar_system_v2
    system = solar_system_v2(moons=True)
    assert len(system) == 26 and len(solar_system_v2()) == 15
    system.standardize_units(mass_unit="kilograms", distance_unit="meters", angle_unit="radians", time_unit="seconds")
    assert [b.name for b in system] == list(g["names"])
    for k, body in enumerate(system):
        r, v = body.get_state()
        assert r == list(g["r"][k]) and v == list(g["v"][k]), body.name
        assert body.mu == g["mu"][k] and body.fg == g["fg"][k] and body.b.value == g["b"][k]
        assert body.mass.value == g["mass"][k] and body.radius.value == g["radius"][k]
        if body.parent is None:
            assert body.T is None and body.mean_motion() == 0.0
        else:
            assert body.T.value == g["T"][k]
    assert system[3].name == "Earth" and system[3].parent is system[0]
    assert set(system.to_json()["Earth"]) >= {"name", "a", "e", "I", "L", "M", "mass", "radius", "mu", "fg", "T", "parent"}


def test_solve_kepler_bit_exact(golden):
    from core.physics import solve_kepler
    g = golden("kepler")
    for i, e in enumerate(g["kep_e"]):
        for j, M in enumerate(g["kep_M"]):
            assert solve_kepler(float(M), float(e)) == g["kep_E"][i, j]


def test_units_and_constants():
    from core import constants as c
    from core import units as u
    assert u.AU_METERS == 1.495978707e11 and u.KG_SOLAR == 1.98847e30
    assert u.Degrees(370).value == 10.0 and u.Degrees(-10).value == 350.0
    assert u.Radians(7.0).value == 7.0 % (2 * math.pi)
    assert (u.Degrees(10) - u.Degrees(20)).value == 350.0
    assert isinstance(u.Degrees(10) + u.Degrees(20), u.Degrees)
    assert u.AU(2).to_meters().value == 2 * 1.495978707e11 and u.Meters(1.0).to_au().unit == "au"
    assert u.Days(2).to_seconds().value == 172800.0 and u.Seconds(43200).to_days().value == 0.5
    assert u.SolarMasses(1).to_kilograms().value == 1.98847e30
    assert repr(u.Meters(3)) == "METERS(3.0)"
    with pytest.raises(ValueError):
        u.Meters(1) + u.AU(1)
    assert c.STANDARD.G == 6.67430e-11 and c.ASTRO.G == 0.0002959122082855911
    assert c.get_unit_profile("SI") is c.STANDARD and c.get_unit_profile(c.UnitSystem.ASTRO) is c.ASTRO
    with pytest.raises(ValueError):
        c.get_unit_profile("cgs")
    with pytest.raises(Exception):
        c.STANDARD.G = 1.0
    assert c.J2000_JD == 2451545.0 and c.JULIAN_DAY == 86400.0 and c.AU == 1.495978707e11 and c.DAY == 86400.0
    assert c.DEFAULT_STANDARD_INTEGRATOR.dt == 3600 and c.DEFAULT_ASTRO_INTEGRATOR.softening == 1e-6


def test_object_semantics():
    from core.physics import Coordinates, Object, ObjectCollection, collide_spheres, moment_of_inertia
    o = Object(2.0, 3.0, velocity=np.array([1.0, 2.0, 3.0]), coordinates=Coordinates(0, 0, 0))
    assert o.velocity.dtype == np.float32 and o.angular_velocity.dtype == np.float32
    assert o.moi == (2 / 5) * 2.0 * 9.0 and o.name == o.uuid[:6]
    assert Object(1.0, 1.0, None, Coordinates(0, 0, 0)).velocity.tolist() == [0, 0, 0]
    assert bool(Coordinates(0, 0, 0))
    d = o.to_dict()
    assert "name" not in d and d["unit_profile"] == "si"
    o2 = Object.from_dict(d)
    assert o2 == o and o2.uuid == o.uuid
    assert moment_of_inertia(2, 3, shape="cylinder") == 9.0 and moment_of_inertia(12, 0, 2, "rod") == 4.0
    with pytest.raises(ValueError):
        moment_of_inertia(1, 1, shape="rod")
    with pytest.raises(ValueError):
        moment_of_inertia(1, 1, shape="cube")
    col = ObjectCollection([o])
    col.append(o2); col.extend([Object(1.0, 1.0, None, Coordinates(5, 0, 0))])
    assert len(col) == 3 and col[0] is o and col.pop() is not None
    col.remove(o2)            # removes by uuid equality -> the first equal element
    assert len(col) == 1
    # separating pair: no-op
    a = Object(1.0, 1.0, np.array([-1.0, 0, 0]), Coordinates(0.0, 0.0, 0.0))
    b = Object(1.0, 1.0, np.array([1.0, 0, 0]), Coordinates(1.0, 0.0, 0.0))
    collide_spheres(a, b)
    assert a.velocity.tolist() == [-1, 0, 0] and a.coordinates == Coordinates(0.0, 0.0, 0.0)


def test_resolve_contacts_equals_full_sweep():
    """The engine's event-driven contact resolution visits the same pairs, in the same order, as the full sweep."""
    from core.physics import Coordinates, Object, ObjectCollection
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = 18
        P = rng.uniform(-3e4, 3e4, (n, 3)); V = rng.standard_normal((n, 3)) * 400
        M = np.exp(rng.uniform(30, 36, n)); R = rng.uniform(2e3, 1.2e4, n)

        def make():
            objs = []
            for i in range(n):
                o = Object(float(M[i]), float(R[i]), V[i], Coordinates(*map(float, P[i])), angular_velocity=np.zeros(3))
                if i % 3 == 0:
                    o.velocity = V[i].copy()
                objs.append(o)
            return ObjectCollection(objs)

        full, fast = make(), make()
        flagged = [(i, j) for i in range(n) for j in range(i + 1, n)
                   if np.linalg.norm(P[i] - P[j]) <= R[i] + R[j]]
        full.handle_collisions(restitution=0.7)
        hits = fast.resolve_contacts(rng.permutation(np.array(flagged).reshape(-1, 2)), restitution=0.7)
        assert hits >= 1
        for a, b in zip(full, fast):
            assert a.position().tolist() == b.position().tolist()
            assert a.velocity.tolist() == b.velocity.tolist() and a.velocity.dtype == b.velocity.dtype


def test_synthetic_ics_are_deterministic():
    from core import synthetic
    d1, d2 = synthetic.uniform_disk(256), synthetic.uniform_disk(256)
    assert np.array_equal(d1["x"], d2["x"]) and d1.n == 256 and d1["m"][0] == synthetic.M_SUN
    p = synthetic.plummer(4096)
    assert abs(np.sum(p["m"] * p["vx"])) < 1e-6 * np.sum(p["m"] * np.abs(p["vx"]))
    assert p["radius"].max() == 0.0 and p["eps"] == 1e9
    e = synthetic.ensemble(3, 16)
    assert e["x"].shape == (3, 16) and np.array_equal(e["x"][1], synthetic.planetary_system(1)["x"])
    assert synthetic.ensemble_fast(8)["m"].shape == (8, 16)
