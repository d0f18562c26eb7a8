#!/usr/bin/env python
"""Time the UNMODIFIED reference (trevormcguire/orbital-physics) on this host's CPU -- run as a subprocess.

    python tools/ref_timing.py --ref baseline/_ref --case solar --n 15 --steps 2000 --vel f32
    python tools/ref_timing.py --ref baseline/_ref --case pairs --n 256
    python tools/ref_timing.py --ref baseline/_ref --case kepler --n 512

`--ref` is a directory holding the reference's own `core/` package (baseline/_ref, a git-ignored copy made by
__graft_entry__.build() from /root/reference; nothing of it is tracked here).  This process imports `core` from
THERE and nothing from this repository, so what is timed is the reference's public API and stock code path:
`SimulationEngine.step` (core/engine.py:65-97) on the solar-system bodies built exactly as
core/examples.py:198-217 builds them (BASELINE config C0), or `pairwise_accelerations` /
`ObjectCollection.handle_collisions` on a random cloud (per-pair cost, used to label extrapolations), or
`Body.get_state` (core/body.py:184-249) on random orbital elements (the initial-condition pipeline).
One JSON object on stdout; the final state is included so that the caller can hold the GPU run to it.
"""
import argparse
import json
import os
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--case", default="solar", choices=["solar", "pairs", "kepler"])
    ap.add_argument("--n", type=int, default=15)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--vel", default="f32", choices=["f32", "f64"])
    a = ap.parse_args()
    ref = os.path.abspath(a.ref)
    if not os.path.isdir(os.path.join(ref, "core")):
        print(json.dumps({"unavailable": f"no reference checkout under {ref}"}))
        return
    sys.path.insert(0, ref)
    import numpy as np
    from core.engine import SimulationEngine          # the reference's, by construction of sys.path
    from core.physics import Coordinates, Object, ObjectCollection, pairwise_accelerations
    import core
    assert os.path.abspath(os.path.dirname(core.__file__)) == os.path.join(ref, "core"), core.__file__
    out = {"case": a.case, "n": a.n, "numpy": np.__version__, "cores_used": 1, "host_cpus": os.cpu_count(),
           "reference_module": os.path.relpath(core.__file__)}

    if a.case == "solar":
        from core.datasets import solar_system_v2
        system = solar_system_v2(moons=False)
        system.standardize_units(mass_unit="kilograms", distance_unit="meters", angle_unit="radians",
                                 time_unit="seconds")
        bodies = []
        for body in list(system)[: a.n]:
            r, v = body.get_state()
            o = Object(mass=body.mass.value, radius=body.radius.value, velocity=np.array(v, dtype=np.float64),
                       coordinates=Coordinates(*r), name=body.name)
            if a.vel == "f64":                          # reassignment keeps float64 (core/physics.py:184 casts only
                o.velocity = np.array(v, dtype=np.float64)   # in the constructor)
            bodies.append(o)
        n = len(bodies)
        eng = SimulationEngine(ObjectCollection(bodies), dt=86400.0, softening=1e6, restitution=1.0, cache=False,
                               max_hist=None)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.step()
        dt = time.perf_counter() - t0
        out.update(n=n, steps=a.steps, velocity_dtype=str(bodies[0].velocity.dtype), seconds=dt,
                   us_per_step=1e6 * dt / a.steps, ordered_interactions_per_s=n * (n - 1) * a.steps / dt,
                   pos=[[float(c) for c in o.position()] for o in bodies],
                   vel=[[float(c) for c in o.velocity] for o in bodies])
    elif a.case == "kepler":
        import core.body as rbody
        import core.units as runits
        rng = np.random.default_rng(a.n)
        sun = rbody.Body(name="Sun", a=runits.Meters(0.0), e=0.0, I=runits.Radians(0.0), L=None, M=runits.Radians(0.0),
                         long_peri=None, long_node=runits.Radians(0.0), arg_peri=runits.Radians(0.0),
                         mass=runits.Kilograms(1.98847e30), radius=runits.Meters(6.9634e8))
        bodies = [rbody.Body(name=f"p{k}", a=runits.Meters(float(np.exp(rng.uniform(np.log(0.1), np.log(40.0))) * 1.495978707e11)),
                             e=float(rng.uniform(0.0, 0.95)), I=runits.Radians(float(abs(rng.normal(0.0, 0.3)))), L=None,
                             M=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))), long_peri=None,
                             long_node=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))),
                             arg_peri=runits.Radians(float(rng.uniform(0.0, 2 * np.pi))),
                             mass=runits.Kilograms(5.9722e24), radius=runits.Meters(6.371e6), parent=sun)
                  for k in range(a.n)]
        t0 = time.perf_counter()
        states = [b.get_state() for b in bodies]
        dt = time.perf_counter() - t0
        el = {"M": [b.M.value for b in bodies], "e": [b.e for b in bodies], "a": [b.a.value for b in bodies],
              "b": [b.b.value for b in bodies], "n": [b.mean_motion() for b in bodies],
              "inc": [b.I.value for b in bodies], "Omega": [b.long_node.value for b in bodies],
              "omega": [b.arg_peri.value for b in bodies]}
        out.update(us_per_body=1e6 * dt / a.n, seconds=dt, elements={k: [float(x) for x in v] for k, v in el.items()},
                   r=[[float(c) for c in s[0]] for s in states], v=[[float(c) for c in s[1]] for s in states])
    else:
        rng = np.random.default_rng(a.n)
        objs = [Object(mass=float(m), radius=1.0, velocity=v, coordinates=Coordinates(*p))
                for m, v, p in zip(np.exp(rng.uniform(np.log(1e20), np.log(1e26), a.n)),
                                   rng.standard_normal((a.n, 3)) * 1e3, rng.uniform(-1e11, 1e11, (a.n, 3)))]
        pairs = a.n * (a.n - 1) // 2
        t0 = time.perf_counter()
        pairwise_accelerations(objs, eps=1e6)
        t_force = time.perf_counter() - t0
        coll = ObjectCollection(objs)
        t0 = time.perf_counter()
        coll.handle_collisions(restitution=1.0)
        t_coll = time.perf_counter() - t0
        out.update(pairs=pairs, force_us_per_pair=1e6 * t_force / pairs, collisions_us_per_pair=1e6 * t_coll / pairs,
                   ordered_interactions_per_s=2 * pairs / (t_force + t_coll))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
