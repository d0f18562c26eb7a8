#!/usr/bin/env python
"""Sweep the fast force kernel's launch geometry on one GPU (development aid, not a bench contract).

    python tools/sweep_fast.py [N] [reps]

For each targets-per-thread (ORBITAL_B200_TI) and a few slab counts (ORBITAL_B200_SLABS) times orb_accel
with CUDA events on the launching stream and prints interactions/s and the 20-flop TFLOP/s equivalent.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))

import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    c = synthetic.plummer(n)
    peak = _native.fp64_peak(0, 0.5)
    print(f"fp64 peak: best {peak['tflops_best']:.2f} mean {peak['tflops_mean']:.2f} TF, clock {peak['sm_clock_mhz']:.0f} MHz")
    stream = torch.cuda.current_stream().cuda_stream
    mode = sys.argv[3] if len(sys.argv) > 3 else "sym"
    if mode == "sym":
        configs = [(ti, ch) for ti in (1, 2, 4, 6) for ch in (None,)] + [(4, ch) for ch in (4, 8, 16, 32, 64, 96)] \
            + [(6, ch) for ch in (16, 48)]
    else:
        configs = [(ti, s) for ti in (1, 2, 4, 6, 8) for s in (None,)] + [(8, s) for s in (1, 4, 8, 16)]
    for ti, slabs in configs:
        os.environ["ORBITAL_B200_SYM"] = "1" if mode == "sym" else "0"
        os.environ["ORBITAL_B200_TI"] = str(ti)
        os.environ["ORBITAL_B200_SYM_TI"] = str(ti)
        for key in ("ORBITAL_B200_SLABS", "ORBITAL_B200_SYM_CHUNKS"):
            if slabs is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = str(slabs)
        dev = _native.DeviceSystem(n, 0, _native.MODE_FAST)
        dev.set_stream(stream)
        dev.set_params(c["dt"], c["eps"], c["G"])
        dev.upload(*c.arrays())
        info = dev.force_kernel_info()
        dev.accel()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dev.accel(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        rate = float(n) * n / (ms * 1e-3)
        print(f"TI={ti} slabs={slabs} grid={info['grid']:6d} launches/step={info['launches_per_step']} "
              f"{ms:8.3f} ms  {rate:.3e} int/s  {20 * rate / 1e12:6.2f} TF  frac(mean peak)={20 * rate / 1e12 / peak['tflops_mean']:.3f}",
              flush=True)
        dev.close()


if __name__ == "__main__":
    main()
