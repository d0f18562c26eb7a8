#!/usr/bin/env python
"""A few bit-exact or fast leapfrog steps at size N (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = _native.MODE_FAST if (len(sys.argv) > 2 and sys.argv[2] == "fast") else _native.MODE_FAITHFUL
c = synthetic.random_cloud(n, seed=n) if mode == _native.MODE_FAITHFUL else synthetic.plummer(n)
dev = _native.DeviceSystem(n, 0, mode)
dev.set_params(c["dt"], c["eps"], c["G"])
dev.upload(*c.arrays())
dev.accel()
dev.step(4)
dev.step(4)
dev.close()
