#!/usr/bin/env python
"""Tiny driver for profiling the bit-exact force pass: N bodies (default 4096), 3 passes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
from core import _native, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = synthetic.uniform_disk(n)
dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays())
for _ in range(3):
    dev.accel()
dev.synchronize()
print("ok", dev.force_kernel_info())
