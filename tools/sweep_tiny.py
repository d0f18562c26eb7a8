#!/usr/bin/env python
"""Bit-exact stepping rate vs size: fused single-CTA kernels against the multi-CTA kernel sequence (development aid).

Run once per ORBITAL_B200_TINY_MAX setting (the limit is read once per process):
    ORBITAL_B200_TINY_MAX=512 python tools/sweep_tiny.py ; ORBITAL_B200_TINY_MAX=64 python tools/sweep_tiny.py
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402

for n in [int(v) for v in os.environ.get("SWEEP_NS", "65,96,128,192,256,384,512").split(",")]:
    c = synthetic.random_cloud(n, seed=n)
    dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
    dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel()
    dev.step(64); dev.synchronize()
    steps = 2048
    t0 = time.perf_counter(); dev.step(steps); dev.synchronize(); dt = time.perf_counter() - t0
    print(f"TINY_MAX={os.environ.get('ORBITAL_B200_TINY_MAX', 'default'):>7s} n={n:4d} {dev.force_kernel_info()['name']:44s} "
          f"{dt / steps * 1e6:8.2f} us/step", flush=True)
    dev.close()
