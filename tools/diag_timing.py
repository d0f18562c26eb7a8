#!/usr/bin/env python
"""Development aid: cost of the on-demand diagnostics (potential, K/L reductions) at bench size."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
from core import _native, synthetic
for n in (262144,):
    c = synthetic.plummer(n)
    dev = _native.DeviceSystem(n, 0, _native.MODE_FAST); dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel(); dev.synchronize()
    for name, fn in (("potential", dev.potential), ("energy_angmom", dev.energy_angmom)):
        fn(); t0 = time.perf_counter(); r = fn(); print(f"n={n} {name}: {(time.perf_counter()-t0)*1e3:.2f} ms -> {r}")
    dev.close()
