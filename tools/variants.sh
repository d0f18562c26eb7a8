#!/bin/bash
# Development aid: time the symmetric kernel for several library builds / TI values on one GPU.
#   tools/variants.sh "liborbital_b200.so liborbital_b200_vmb3.so" "6 8" "1 0"
cd "$(dirname "$0")/.."
LIBS=${1:-liborbital_b200.so}
TIS=${2:-"4 6 8"}
UNIS=${3:-"1 0"}
for lib in $LIBS; do
  for ti in $TIS; do
    for uni in $UNIS; do
    ORBITAL_B200_LIB=$PWD/orbital-physics_b200/csrc/$lib ORBITAL_B200_SYM_TI=$ti ORBITAL_B200_SYM_UNI=$uni python - <<PY
import os, sys
sys.path.insert(0, "orbital-physics_b200")
import numpy as np, torch
from core import _native, synthetic
n = int(os.environ.get("VARIANT_N", "262144"))
c = synthetic.plummer(n)
dev = _native.DeviceSystem(n, 0, _native.MODE_FAST)
dev.set_stream(torch.cuda.current_stream().cuda_stream)
dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel(); torch.cuda.synchronize()
ts = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.accel(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts)); i = dev.force_kernel_info()
print(f"$lib {i['name']} grid={i['grid']} {ms:.3f} ms {n*n/ms/1e9:.4f}e12 int/s", flush=True)
PY
    done
  done
done
