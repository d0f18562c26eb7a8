#!/usr/bin/env python
"""Development aid: micro_steps_kernel per-step time with/without overlap detection and history (n = 15)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
from core import _native, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
c = synthetic.random_cloud(n, seed=3)
for radius in (0.0, 1.0):
    for hist in (0, 64):
        arrs = list(c.arrays()); arrs[7] = np.full(n, radius)
        dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
        dev.set_params(c["dt"], c["eps"], c["G"])
        if hist:
            dev.set_history(hist)
        dev.upload(*arrs); dev.accel(); dev.step(100); dev.synchronize()
        t0 = time.perf_counter(); dev.step(20000); dev.synchronize(); dt = time.perf_counter() - t0
        print(f"n={n} radius={radius:g} history={hist}: {dt / 20000 * 1e6:.3f} us/step", flush=True)
        dev.close()
