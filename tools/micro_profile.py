#!/usr/bin/env python
"""Tiny driver for profiling the micro-system kernel: solar system N=15, one launch of 20,000 steps."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200")); sys.path.insert(0, REPO)
from core.engine import SimulationEngine
from core.examples import solar_system_objects
from core.physics import ObjectCollection
bodies, _ = solar_system_objects(moons=False)
eng = SimulationEngine(ObjectCollection(bodies), dt=86400.0, softening=1e6, cache=False, max_hist=None)
eng.run(20000)
eng.synchronize()
print("ok", eng.kernel_info())
