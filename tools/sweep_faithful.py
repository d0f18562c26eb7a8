#!/usr/bin/env python
"""Faithful (bit-exact) force pass vs targets-per-warp TW on one GPU (development aid)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402

for n in (1024, 2048, 4096, 8192, 16384):
    c = synthetic.uniform_disk(n) if hasattr(synthetic, "uniform_disk") else synthetic.random_cloud(n, seed=n)
    row = [f"N={n:6d}"]
    ref = None
    for tw, pairs in ((1, "0"), (2, "0"), (4, "0"), (8, "0"), (0, "1")):
        os.environ["ORBITAL_B200_FAITHFUL_TW"] = str(tw)
        os.environ["ORBITAL_B200_FAITHFUL_PAIRS"] = pairs
        dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
        dev.set_stream(torch.cuda.current_stream().cuda_stream)
        dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel(); dev.accel()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dev.accel(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        acc = dev.download_acc()
        if ref is None:
            ref = acc
        same = np.array_equal(acc, ref)
        row.append(f"{'2pass ' if pairs == '1' else ''}TW={tw}: {np.median(ts):7.4f}{'' if same else ' MISMATCH'}")
        dev.close()
    print("  ".join(row), flush=True)
