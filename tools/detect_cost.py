#!/usr/bin/env python
"""Development aid: cost of the fused overlap detection in the fast force pass (radii > 0) at bench size."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
import torch
from core import _native, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
c = synthetic.plummer(n)
for radius in (0.0, 1.0e3):
    arrs = list(c.arrays()); arrs[7] = np.full(n, radius)
    dev = _native.DeviceSystem(n, 0, _native.MODE_FAST)
    dev.set_stream(torch.cuda.current_stream().cuda_stream)
    dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*arrs); dev.accel(); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dev.accel(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"radius={radius:g}: {dev.force_kernel_info()['name']} {np.median(ts):.3f} ms", flush=True)
    dev.close()
