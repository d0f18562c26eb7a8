#!/usr/bin/env python
"""Fast-mode leapfrog step (orb_step under its CUDA graph) vs problem size, TI and source-tile size (GPU box).

    python tools/sweep_step.py [uniform|random] [N ...]  > profiles/rN_sweep_step.txt

us per step and interactions/s (N^2 per step) for the default geometry and for every forced (TI, tile); the
heuristics in plan_sym (csrc/force_sym.cu) are read off this table."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402

PEAK = 1.86e12      # interactions/s at the 37.2 TF nominal FP64 peak, 20 flop each


def step_us(c, steps):
    dev = _native.DeviceSystem(c.n, 0, _native.MODE_FAST)
    dev.set_params(c["dt"], c["eps"], c["G"])
    dev.upload(*c.arrays())
    dev.accel()
    dev.step(32)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        dev.step(steps)
        best = min(best, time.perf_counter() - t0)
    name = dev.force_kernel_info()["name"]
    dev.close()
    return 1e6 * best / steps, name


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    sizes = [int(a) for a in sys.argv[2:]] or [1024, 2048, 4096, 8192, 16384, 32768, 65536]
    for n in sizes:
        c = synthetic.plummer(n) if kind == "uniform" else synthetic.random_cloud(n, seed=n)
        steps = 512 if n <= 8192 else (128 if n <= 32768 else 32)
        for k in ("ORBITAL_B200_SYM_TI", "ORBITAL_B200_SYM_TILE"):
            os.environ.pop(k, None)
        us, name = step_us(c, steps)
        print(f"N={n:6d} default {name:30s} {us:9.2f} us/step {n * n / us / 1e6:7.4f}e12 int/s "
              f"({100 * n * n / us * 1e6 / PEAK:5.1f} % of nominal)", flush=True)
        for ti in (1, 2, 4, 8):
            row = []
            for tile in (64, 128, 256):
                os.environ["ORBITAL_B200_SYM_TI"] = str(ti)
                os.environ["ORBITAL_B200_SYM_TILE"] = str(tile)
                u, _ = step_us(c, max(16, steps // 4))
                row.append(f"tile {tile:3d}: {u:9.2f}")
            print(f"          TI={ti}  " + "   ".join(row), flush=True)


if __name__ == "__main__":
    main()
