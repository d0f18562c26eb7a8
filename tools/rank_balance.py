#!/usr/bin/env python
"""Per-rank force time of a `world`-rank fast-mode job, measured by running every rank's kernels on ONE GPU
(LocalComm([0] * world)): max / mean over ranks is the load balance of the pair-block ownership.

    python tools/rank_balance.py [N] [world]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402
from core.distributed import LocalComm, ShardedSystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
c = synthetic.plummer(n)
sh = ShardedSystem.from_arrays(*c.arrays(), c["dt"], c["eps"], c["G"], mode=_native.MODE_FAST, comm=LocalComm([0] * world))
sh.step(1)
times = []
for d in sh.devs:
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); d.step_force(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    times.append(best)
t = np.array(times)
print(f"N={n} world={world}: force ms per rank {np.round(t, 3).tolist()}")
print(f"max {t.max():.3f} mean {t.mean():.3f} -> balance {t.mean() / t.max():.4f}; sum {t.sum():.2f} ms (one GPU, whole pass ~{39.4 * (n / 262144) ** 2:.1f})")
sh.close()
