#!/usr/bin/env python
"""Ensemble throughput by batch size (GPU box):  python tools/ens_sweep.py [nbody] > profiles/rN_ens_sizes.txt

One step per launch (HBM form: x,u,m in / x,u out = 104 B per body-step, +48 B per 16 launches) and all steps fused
in one launch, for the batch sizes a 1 / 2 / 4 / 8-GPU split of BASELINE configs[3] gives each GPU, and two batches
beyond the L2."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "orbital-physics_b200"), REPO):
    sys.path.insert(0, p)
import torch  # noqa: E402
from core import _native, synthetic  # noqa: E402


def main():
    nbody = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = 208
    print(f"# nbody={nbody}, {steps} steps; GB/s on the moved bytes (107 B/body-step) and on SURVEY's 104 B")
    print(f"{'systems':>9s} {'unfused us/step':>16s} {'GB/s moved':>11s} {'GB/s 104B':>10s} {'fused us/step':>14s} {'fused int/s':>12s}")
    for nsys in (8192, 16384, 32768, 65536, 262144, 524288):
        e = synthetic.ensemble_fast(nsys, nbody)
        ens = _native.DeviceEnsemble(nsys, nbody, 0, _native.MODE_FAST)
        ens.set_stream(torch.cuda.current_stream().cuda_stream)
        ens.set_params(e["dt"], e["eps"], e["G"])
        ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
        out = []
        for fused in (False, True):
            ens.step(32, fused=fused)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ens.step(steps, fused=fused); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            out.append(best)
        moved = ens.info()["bytes_per_body_step"]
        ens.close()
        bodies = nsys * nbody
        print(f"{nsys:9d} {1e3 * out[0] / steps:16.2f} {moved * bodies * steps / (out[0] * 1e-3) / 1e9:11.0f} "
              f"{104.0 * bodies * steps / (out[0] * 1e-3) / 1e9:10.0f} {1e3 * out[1] / steps:14.2f} "
              f"{bodies * nbody * steps / (out[1] * 1e-3):12.4g}")


if __name__ == "__main__":
    main()
