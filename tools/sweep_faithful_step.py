#!/usr/bin/env python
"""Bit-exact leapfrog step (orb_step under its CUDA graph) vs problem size, both pass-2 kernels (GPU box).

    python tools/sweep_faithful_step.py > profiles/rN_sweep_faithful_step.txt"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402


def step_us(c, steps):
    dev = _native.DeviceSystem(c.n, 0, _native.MODE_FAITHFUL)
    dev.set_params(c["dt"], c["eps"], c["G"])
    dev.upload(*c.arrays())
    dev.accel()
    dev.step(16)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        dev.step(steps)
        best = min(best, time.perf_counter() - t0)
    st = dev.download_state()
    dev.close()
    return 1e6 * best / steps, st


which = {"1": "one warp per 32 targets", "2": "default: producer / consumer warps up to 148 blocks, one warp per 32 targets above",
         "4": "four lanes per target"}[os.environ.get("ORBITAL_B200_ROWS", "2")]
print(f"# pass 2 of the bit-exact force: {which} (ORBITAL_B200_ROWS = 4 | 2 | 1 selects; read once per process)")
print(f"{'N':>6s} {'us/step':>10s} {'interactions/s':>15s}")
for n in (65, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
    c = synthetic.random_cloud(n, seed=n)
    steps = 256 if n <= 2048 else (64 if n <= 8192 else 16)
    us, _ = step_us(c, steps)
    print(f"{n:6d} {us:10.2f} {n * n / us * 1e6:15.4g}", flush=True)
