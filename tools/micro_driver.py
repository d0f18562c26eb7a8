#!/usr/bin/env python
"""Solar system (N=15) in bit-exact mode: two 2,000-step launches of micro_steps_kernel (for ncu)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core.engine import SimulationEngine  # noqa: E402
from core.examples import solar_system_objects  # noqa: E402
from core.physics import ObjectCollection  # noqa: E402

bodies, _ = solar_system_objects(moons=False)
eng = SimulationEngine(ObjectCollection(bodies), dt=86400.0, softening=1e6, cache=False, max_hist=None)
eng.run(2000)
eng.run(2000)
eng.synchronize()
eng.close()
