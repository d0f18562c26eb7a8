#!/usr/bin/env python
"""Time the UNMODIFIED reference (trevormcguire/orbital-physics) on this host's CPU -- run as a subprocess.

    python tools/ref_timing.py --ref baseline/_ref --case solar --n 15 --steps 2000 --vel f32
    python tools/ref_timing.py --ref baseline/_ref --case pairs --n 256

`--ref` is a directory holding the reference's own `core/` package (baseline/_ref, a git-ignored copy made by
__graft_entry__.build() from /root/reference; nothing of it is tracked here).  This process imports `core` from
THERE and nothing from this repository, so what is timed is the reference's public API and stock code path:
`SimulationEngine.step` (core/engine.py:65-97) on the solar-system bodies built exactly as
core/examples.py:198-217 builds them (BASELINE config C0), or `pairwise_accelerations` /
`ObjectCollection.handle_collisions` on a random cloud (per-pair cost, used to label extrapolations).
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
    ap.add_argument("--case", default="solar", choices=["solar", "pairs"])
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
