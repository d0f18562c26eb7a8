#!/bin/bash
# Development aid: time the symmetric kernel for several library builds / TI values on one GPU.
cd "$(dirname "$0")/.."
for lib in liborbital_b200.so; do
  for ti in 4 6 8; do
    echo "== $lib TI=$ti"
    ORBITAL_B200_LIB=$PWD/orbital-physics_b200/csrc/$lib ORBITAL_B200_SYM_TI=$ti python - <<PY
import os, sys
sys.path.insert(0, "orbital-physics_b200")
import numpy as np, torch
from core import _native, synthetic
n = 262144
c = synthetic.plummer(n)
dev = _native.DeviceSystem(n, 0, _native.MODE_FAST)
dev.set_stream(torch.cuda.current_stream().cuda_stream)
dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.accel(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts)); print(f"{dev.force_kernel_info()['name']} {ms:.3f} ms {n*n/ms/1e9:.4f}e12 int/s")
PY
  done
done
