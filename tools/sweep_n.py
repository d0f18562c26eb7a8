#!/usr/bin/env python
"""Fast-mode force pass vs problem size on one GPU (development aid): for each N, time orb_accel with the default
TI heuristic and with every forced TI, so the heuristic in plan_sym (csrc/force_sym.cu) can be checked.

    python tools/sweep_n.py [uniform|random]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))

import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402


def time_accel(c, reps=5):
    dev = _native.DeviceSystem(c.n, 0, _native.MODE_FAST)
    dev.set_stream(torch.cuda.current_stream().cuda_stream)
    dev.set_params(c["dt"], c["eps"], c["G"])
    dev.upload(*c.arrays())
    dev.accel(); dev.accel()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dev.accel(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    name = dev.force_kernel_info()["name"]
    dev.close()
    return float(np.median(ts)), name


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
        c = synthetic.plummer(n) if kind == "uniform" else synthetic.random_cloud(n, seed=n)
        os.environ.pop("ORBITAL_B200_SYM_TI", None)
        ms, name = time_accel(c)
        row = [f"N={n:7d} default {name:32s} {ms:9.4f} ms {n * n / ms / 1e9:8.4f}e12 int/s |"]
        for ti in (1, 2, 4, 8):
            os.environ["ORBITAL_B200_SYM_TI"] = str(ti)
            ms_t, _ = time_accel(c, reps=3)
            row.append(f"TI={ti}: {ms_t:8.4f}")
        print(" ".join(row), flush=True)


if __name__ == "__main__":
    main()
