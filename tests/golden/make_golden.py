#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py [--skip-disk]

It imports the reference's own ``core`` package (matplotlib stubbed: it is not
installed and only the plotting helpers need it), drives
``pairwise_accelerations`` / ``SimulationEngine.step`` / ``handle_collisions``
/ ``Body.get_state`` on fixed inputs and stores inputs + outputs as ``.npz``
fixtures next to this file.  Nothing here is product code and no reference
source is copied: the fixtures are the reference's *outputs*.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ORBITAL_REFERENCE", "/root/reference")

# --- import the reference -------------------------------------------------
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
import core.physics as rphys            # noqa: E402  (the reference)
import core.engine as reng              # noqa: E402
import core.datasets as rdata           # noqa: E402
assert rphys.__file__.startswith(REF), rphys.__file__

# our synthetic IC module, loaded by path (its package is also called "core")
_spec = importlib.util.spec_from_file_location(
    "b200_synthetic", os.path.join(REPO, "orbital-physics_b200", "core", "synthetic.py"))
syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(syn)


def make_objects(x, y, z, vx, vy, vz, m, radius, f64_velocity):
    """Build reference Objects.  f64_velocity[i] => reassign .velocity (fp64 mode)."""
    objs = []
    for i in range(len(x)):
        v = np.array([vx[i], vy[i], vz[i]], dtype=np.float64)
        o = rphys.Object(mass=float(m[i]), radius=float(radius[i]), velocity=v,
                         coordinates=rphys.Coordinates(float(x[i]), float(y[i]), float(z[i])),
                         angular_velocity=np.zeros(3), name=f"b{i}")
        if f64_velocity[i]:
            o.velocity = v.copy()          # like core/examples.py:104-105 / physics.py:448-449
        objs.append(o)
    return objs


def snapshot(engine):
    objs = engine.objects.objects
    pos = np.array([[o.coordinates.x, o.coordinates.y, o.coordinates.z] for o in objs], dtype=np.float64)
    vel = np.array([np.asarray(o.velocity, dtype=np.float64) for o in objs])
    acc = np.array([engine.acc[o.uuid] for o in objs])
    return pos, vel, acc


def run_case(cloud, f64_velocity, steps_to_record, restitution=1.0, dt=None, eps=None):
    """ctor + stepping with the real engine; record state at the listed steps."""
    n = len(cloud["x"])
    flags = np.broadcast_to(np.asarray(f64_velocity, dtype=bool), (n,)).copy()
    objs = make_objects(*(cloud[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius")), flags)
    dt = cloud["dt"] if dt is None else dt
    eps = cloud["eps"] if eps is None else eps
    eng = reng.SimulationEngine(rphys.ObjectCollection(objs), dt=dt, softening=eps,
                                restitution=restitution, max_hist=None, cache=False)
    out = {"dt": dt, "eps": eps, "restitution": restitution, "f64_velocity": flags,
           "steps": np.array(sorted(steps_to_record))}
    for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius"):
        out["in_" + k] = np.asarray(cloud[k], dtype=np.float64)
    # the velocities the engine actually starts from (fp32-rounded where applicable)
    p0, v0, a0 = snapshot(eng)
    out["pos_0"], out["vel_0"], out["acc_0"] = p0, v0, a0
    out["U_0"] = float(eng.last_potential)
    out["E_0"] = float(eng.total_energy())
    out["L_0"] = np.asarray(eng.angular_momentum(), dtype=np.float64)
    done = 0
    for s in sorted(steps_to_record):
        while done < s:
            eng.step()
            done += 1
        p, v, a = snapshot(eng)
        out[f"pos_{s}"], out[f"vel_{s}"], out[f"acc_{s}"] = p, v, a
        out[f"U_{s}"] = float(eng.last_potential)
        out[f"E_{s}"] = float(eng.total_energy())
        out[f"L_{s}"] = np.asarray(eng.angular_momentum(), dtype=np.float64)
    return out


def solar_cloud(moons: bool, parent_offset: bool, first: int | None = None):
    """Reference's own Kepler pipeline -> SI state (examples.py:198-217, app.py:27-50)."""
    system = rdata.solar_system_v2(moons=moons)
    system.standardize_units(mass_unit="kilograms", distance_unit="meters",
                             angle_unit="radians", time_unit="seconds")
    rows, names = [], []
    for body in system:
        r, v = body.get_state()
        if parent_offset and body.parent is not None:
            pr, pv = body.parent.get_state()
            r = np.array(pr) + np.array(r)
            v = np.array(pv) + np.array(v)
        rows.append((list(map(float, r)), list(map(float, v)), body.mass.value, body.radius.value))
        names.append(body.name)
    if first is not None:
        rows, names = rows[:first], names[:first]
    pos = np.array([r[0] for r in rows])
    vel = np.array([r[1] for r in rows])
    c = dict(x=pos[:, 0].copy(), y=pos[:, 1].copy(), z=pos[:, 2].copy(),
             vx=vel[:, 0].copy(), vy=vel[:, 1].copy(), vz=vel[:, 2].copy(),
             m=np.array([r[2] for r in rows]), radius=np.array([r[3] for r in rows]))
    return c, names


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz ({os.path.getsize(path) / 1024:.1f} KiB)", flush=True)


def golden_force():
    """pairwise_accelerations on random clouds (core/physics.py:125-159)."""
    out = {}
    cases = [(2, 0.0), (3, 1e7), (9, 1e6), (16, 0.0), (33, 1e8), (64, 1e8), (257, 1e8), (1024, 1e8)]
    for k, (n, eps) in enumerate(cases):
        c = syn.random_cloud(n, seed=100 + n)
        objs = make_objects(*c.arrays(), np.ones(n, bool))
        acc, U = rphys.pairwise_accelerations(objs, eps=eps)
        a = np.array([acc[o.uuid] for o in objs])
        for key in ("x", "y", "z", "m"):
            out[f"c{k}_{key}"] = c[key]
        out[f"c{k}_eps"] = eps
        out[f"c{k}_acc"] = a
        out[f"c{k}_U"] = float(U)
    out["ncases"] = len(cases)
    # a cloud with two coincident bodies and eps > 0 (finite) and eps = 0 (inf/nan, no exception)
    c = syn.random_cloud(5, seed=7)
    c["x"][3], c["y"][3], c["z"][3] = c["x"][1], c["y"][1], c["z"][1]
    for tag, eps in (("coinc_soft", 1e5), ("coinc_hard", 0.0)):
        objs = make_objects(*c.arrays(), np.ones(5, bool))
        with np.errstate(all="ignore"):
            acc, U = rphys.pairwise_accelerations(objs, eps=eps)
        out[f"{tag}_acc"] = np.array([acc[o.uuid] for o in objs])
        out[f"{tag}_U"] = float(U)
        out[f"{tag}_eps"] = eps
    for key in ("x", "y", "z", "m"):
        out[f"coinc_{key}"] = c[key]
    save("force_random", **out)


def golden_ddot():
    """Pin the rounding of NumPy's 3-element ``rij @ rij`` on this host."""
    rng = np.random.default_rng(5)
    v = rng.standard_normal((4000, 3)) * np.exp(rng.uniform(-20, 20, (4000, 1)))
    d = np.array([float(r @ r) for r in v])
    n = np.array([float(np.linalg.norm(r)) for r in v])
    save("ddot3", v=v, dot=d, norm=n)


def golden_solar():
    for tag, moons, offset, first, dt in (("solar15", False, False, None, 86400.0),
                                          ("solar9", False, False, 9, 86400.0),
                                          ("solar26", True, True, None, 1800.0)):
        c, names = solar_cloud(moons, offset, first)
        c.update(dt=dt, eps=1e6)
        steps = [1, 2, 10, 100, 1000] + ([10000] if tag != "solar26" else [])
        for mode, f64 in (("f32", False), ("f64", True)):
            t0 = time.time()
            out = run_case(c, f64, steps)
            out["names"] = np.array(names)
            save(f"{tag}_{mode}", **out)
            print(f"    {tag}_{mode}: {time.time() - t0:.1f}s", flush=True)


def golden_mixed():
    """Mixed velocity dtypes across bodies + a scenario per reference example."""
    c = syn.random_cloud(12, seed=12, scale=2e11, radius=1e3)
    flags = np.arange(12) % 3 == 0
    save("mixed12", **run_case(c, flags, [1, 5, 50, 500]))
    # two-body circular orbit (examples.py:11-49): fp64 velocities from set_circular_orbit
    b1 = rphys.Object(5.972e24, 6.371e6, velocity=np.zeros(3), coordinates=rphys.Coordinates(0, 0, 0))
    b2 = rphys.Object(7.348e22, 1.737e6, velocity=np.zeros(3), coordinates=rphys.Coordinates(384400e3, 0, 0))
    rphys.set_circular_orbit(primary=b1, secondary=b2)
    eng = reng.SimulationEngine(rphys.ObjectCollection([b1, b2]), dt=3600.0, softening=1e3, cache=False, max_hist=None)
    v0 = np.array([b1.velocity, b2.velocity])
    for _ in range(500):
        eng.step()
    p, v, a = snapshot(eng)
    save("two_body", v0=v0, pos_500=p, vel_500=v, acc_500=a, E_500=float(eng.total_energy()),
         L_500=eng.angular_momentum(), hist_b2=np.array(eng.history[b2.uuid]))
    # three-body equilateral (examples.py:124-178): fp32 velocities through the ctor
    m, R = 1e22, 1e7
    pos = [np.array([R, 0.0, 0.0]), np.array([-0.5 * R, np.sqrt(3) / 2 * R, 0.0]),
           np.array([-0.5 * R, -np.sqrt(3) / 2 * R, 0.0])]
    z_hat = np.array([0.0, 0.0, 1.0])
    t_hat = [np.cross(z_hat, p / np.linalg.norm(p)) for p in pos]
    vmag = np.sqrt(rphys.STANDARD.G * m / (np.sqrt(3.0) * R))
    objs = [rphys.Object(mass=m, radius=(m / 5000.0) ** (1 / 3), velocity=vmag * t_hat[i],
                         coordinates=rphys.Coordinates.from_iterable(pos[i]), angular_velocity=np.zeros(3))
            for i in range(3)]
    eng = reng.SimulationEngine(rphys.ObjectCollection(objs), dt=50.0, softening=1e3, cache=False, max_hist=None)
    for _ in range(1000):
        eng.step()
    p, v, a = snapshot(eng)
    save("three_body", pos_1000=p, vel_1000=v, acc_1000=a, E_1000=float(eng.total_energy()),
         radius=objs[0].radius)


def golden_collisions():
    """handle_collisions / collide_spheres (core/physics.py:391-422,510-535) through the engine."""
    # (a) head-on pair + bystander, both velocity modes, restitution 1.0 and 0.5
    for tag, f64, e in (("hit_f32_e1", False, 1.0), ("hit_f64_e1", True, 1.0),
                        ("hit_f32_e05", False, 0.5), ("hit_f64_e05", True, 0.5)):
        c = dict(x=np.array([-1.0e4, 1.0e4, 0.0]), y=np.array([0.0, 30.0, 5.0e6]), z=np.array([0.0, -20.0, 0.0]),
                 vx=np.array([900.0, -700.0, 0.0]), vy=np.array([3.0, 0.0, 1.0]), vz=np.array([0.0, 2.0, 0.0]),
                 m=np.array([4.0e15, 9.0e15, 1.0e12]), radius=np.array([3.0e3, 4.0e3, 10.0]),
                 dt=1.0, eps=1.0)
        save("coll_" + tag, **run_case(c, f64, list(range(1, 21)) + [40, 80], restitution=e))
    # (b) dense cloud with big radii: many overlapping pairs per step, sequential in-place semantics matter
    rng = np.random.default_rng(77)
    n = 24
    c = dict(x=rng.uniform(-4e4, 4e4, n), y=rng.uniform(-4e4, 4e4, n), z=rng.uniform(-4e4, 4e4, n),
             vx=rng.standard_normal(n) * 300, vy=rng.standard_normal(n) * 300, vz=rng.standard_normal(n) * 300,
             m=np.exp(rng.uniform(np.log(1e14), np.log(1e16), n)), radius=rng.uniform(2e3, 8e3, n),
             dt=2.0, eps=10.0)
    save("coll_dense_f32", **run_case(c, False, list(range(1, 31)), restitution=0.8))
    flags = np.arange(n) % 2 == 0
    save("coll_dense_mixed", **run_case(c, flags, list(range(1, 31)), restitution=1.0))


def golden_disk():
    """C1: uniform disk N=4096, ctor + 2 steps of the real engine (~5 min on one core)."""
    c = syn.uniform_disk(4096)
    t0 = time.time()
    out = run_case(c, False, [1, 2])
    out["seconds"] = time.time() - t0
    # inputs are regenerated from the seed by the tests; keep outputs only (fixture size)
    for k in list(out):
        if k.startswith("in_"):
            del out[k]
    save("disk4096_f32", **out)


def golden_kepler():
    """Body.get_state / derive over the dataset (core/body.py:65-97,184-249; datasets.py:13-56)."""
    system = rdata.solar_system_v2(moons=True)
    system.standardize_units(mass_unit="kilograms", distance_unit="meters",
                             angle_unit="radians", time_unit="seconds")
    r, v, mu, fg, T, b, names = [], [], [], [], [], [], []
    for body in system:
        rr, vv = body.get_state()
        r.append(rr); v.append(vv); mu.append(body.mu); fg.append(body.fg)
        T.append(body.T.value if body.T is not None else np.nan)
        b.append(body.b.value); names.append(body.name)
    M = np.linspace(0, 2 * np.pi, 97)
    ecc = np.array([0.0, 0.01, 0.2, 0.5, 0.79, 0.8, 0.9, 0.99])
    E = np.array([[rphys.solve_kepler(float(m_), float(e_)) for m_ in M] for e_ in ecc])
    save("kepler", r=np.array(r), v=np.array(v), mu=np.array(mu), fg=np.array(fg), T=np.array(T),
         b=np.array(b), names=np.array(names), kep_M=M, kep_e=ecc, kep_E=E,
         mass=np.array([bd.mass.value for bd in system]), radius=np.array([bd.radius.value for bd in system]))


def golden_kepler_batch(count=2048, seed=43):
    """Random element sets through the reference's Body.get_state (core/body.py:184-249) -- the fixture the
    batched device pipeline (orb_kepler_states / orb_ens_upload_elements) and its oracle restatement are held to."""
    import core.body as rbody
    import core.units as runits
    rng = np.random.default_rng(seed)
    sun = rbody.Body(name="Sun", a=runits.Meters(0.0), e=0.0, I=runits.Radians(0.0), L=None, M=runits.Radians(0.0),
                     long_peri=None, long_node=runits.Radians(0.0), arg_peri=runits.Radians(0.0),
                     mass=runits.Kilograms(1.98847e30), radius=runits.Meters(6.9634e8))
    ecc = np.concatenate([rng.uniform(0.0, 0.97, count - 8), [0.0, 0.79999, 0.8, 0.8000001, 0.9, 0.95, 0.97, 0.5]])
    out = {k: np.empty(count) for k in ("M", "e", "a", "b", "n", "inc", "Omega", "omega", "E")}
    r, v = np.empty((count, 3)), np.empty((count, 3))
    for k in range(count):
        body = rbody.Body(name=f"p{k}", a=runits.Meters(float(np.exp(rng.uniform(np.log(0.1), np.log(40.0))) * 1.495978707e11)),
                          e=float(ecc[k]), I=runits.Radians(float(abs(rng.normal(0.0, 0.3)))), L=None,
                          M=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))), long_peri=None,
                          long_node=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))),
                          arg_peri=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))),
                          mass=runits.Kilograms(5.9722e24), radius=runits.Meters(6.371e6), parent=sun)
        rr, vv = body.get_state()
        r[k], v[k] = rr, vv
        out["M"][k] = body.M.value; out["e"][k] = body.e; out["a"][k] = body.a.value; out["b"][k] = body.b.value
        out["n"][k] = body.mean_motion(); out["inc"][k] = body.I.value; out["Omega"][k] = body.long_node.value
        out["omega"][k] = body.arg_peri.value
        out["E"][k] = rphys.solve_kepler(body.M.value, body.e)
    save("kepler_batch", r=r, v=v, parent_mass=np.array(sun.mass.value), parent_mu=np.array(sun.mu), **out)


def main():
    skip_disk = "--skip-disk" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    jobs = dict(ddot=golden_ddot, force=golden_force, kepler=golden_kepler, kepler_batch=golden_kepler_batch,
                mixed=golden_mixed,
                collisions=golden_collisions, solar=golden_solar, disk=golden_disk)
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        if name == "disk" and skip_disk:
            continue
        print(f"[{name}]", flush=True)
        fn()


if __name__ == "__main__":
    np.seterr(all="ignore")
    main()
