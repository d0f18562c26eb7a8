#!/usr/bin/env python
"""Un-fused ensemble step time for small batches vs the number of independent graph branches (GPU box)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)
import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402

steps = 416
print(f"{'systems':>8s} {'branches':>9s} {'pdl':>4s} {'us/step':>9s}")
for nsys in [int(a) for a in sys.argv[1:]] or (4096, 8192, 16384, 32768):
    e = synthetic.ensemble_fast(nsys, 16)
    for br in (1, 2, 3, 4):
        for pdl in (1, 0):
            os.environ["ORBITAL_B200_ENS_BRANCHES"] = str(br)
            os.environ["ORBITAL_B200_ENS_PDL"] = str(pdl)
            ens = _native.DeviceEnsemble(nsys, 16, 0, _native.MODE_FAST)
            ens.set_stream(torch.cuda.current_stream().cuda_stream)
            ens.set_params(e["dt"], e["eps"], e["G"])
            ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
            ens.step(steps, fused=False)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ens.step(steps, fused=False); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            ens.close()
            print(f"{nsys:8d} {br:9d} {pdl:4d} {1e3 * best / steps:9.3f}")
