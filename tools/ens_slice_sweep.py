#!/usr/bin/env python
"""Fused ensemble throughput for small batches vs the time-slice length (GPU box)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)
import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402

steps = 416
print(f"{'systems':>8s} {'slice':>6s} {'crew':>5s} {'us/step':>9s} {'int/s':>11s}")
for nsys in (8192, 16384):
    e = synthetic.ensemble_fast(nsys, 16)
    for sl in (0, 8, 16, 32, 52, 104):
        for crew in ((0,) if sl == 0 else (0, 2, 3)):
            os.environ["ORBITAL_B200_ENS_SLICE"] = str(sl)
            os.environ["ORBITAL_B200_ENS_CREW"] = str(crew)
            ens = _native.DeviceEnsemble(nsys, 16, 0, _native.MODE_FAST)
            ens.set_stream(torch.cuda.current_stream().cuda_stream)
            ens.set_params(e["dt"], e["eps"], e["G"])
            ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
            ens.step(steps, fused=True)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ens.step(steps, fused=True); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            ens.close()
            print(f"{nsys:8d} {sl:6d} {crew:5d} {1e3 * best / steps:9.3f} {nsys * 256 * steps / (best * 1e-3):11.4g}")
