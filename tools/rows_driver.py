#!/usr/bin/env python
"""Tiny driver for profiling the bit-exact force at N=4096: a few orb_accel calls (ncu -k regex:faithful_rows)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c = synthetic.random_cloud(n, seed=n)
dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
dev.set_params(c["dt"], c["eps"], c["G"])
dev.upload(*c.arrays())
for _ in range(3):
    dev.accel()
dev.synchronize()
dev.close()
