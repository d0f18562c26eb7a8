#!/usr/bin/env python
"""Cost of the reference's driver pattern -- `for _ in range(k): engine.step()` -- through SimulationEngine (GPU box).

Solar-system configs (BASELINE configs[0]); compares deferred single steps (default), ORBITAL_B200_DEFER=0 (every
step() is a launch + a device round trip) and run(k).    python tools/step_api_cost.py > profiles/rN_step_api_cost.txt"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)
import core.engine as ce  # noqa: E402
from core.examples import solar_system_objects  # noqa: E402
from core.physics import ObjectCollection  # noqa: E402


def engine(moons):
    bodies, _ = solar_system_objects(moons=moons)
    return ce.SimulationEngine(ObjectCollection(bodies), dt=3600.0, cache=False, max_hist=1)


def timed(moons, k, how):
    eng = engine(moons)
    eng.run(8)
    p0 = eng.objects.objects[1].coordinates.x
    t0 = time.perf_counter()
    if how == "run":
        eng.run(k)
    else:
        for _ in range(k):
            eng.step()
    x = eng.objects.objects[1].coordinates.x            # the read that observes the state
    dt = time.perf_counter() - t0
    n = len(eng.objects.objects)
    eng.close()
    return n, 1e6 * dt / k, x


print(f"{'bodies':>6s} {'pattern':>34s} {'us/step':>9s}   x[1] after the run")
for moons in (False, True):
    for how, defer in (("step() x k, deferred (default)", True), ("step() x k, ORBITAL_B200_DEFER=0", False), ("run", True)):
        ce._DEFER = defer
        n, us, x = timed(moons, 10000, "run" if how == "run" else "step")
        print(f"{n:6d} {how + (' (k)' if how == 'run' else ''):>34s} {us:9.2f}   {x!r}")
