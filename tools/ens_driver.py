#!/usr/bin/env python
"""A few ensemble launches for ncu:  python tools/ens_driver.py <nsys> <fused|unfused> [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200"))
from core import _native, synthetic  # noqa: E402

nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fused = (sys.argv[2] if len(sys.argv) > 2 else "fused") == "fused"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else (64 if fused else 18)
e = synthetic.ensemble_fast(nsys, 16)
ens = _native.DeviceEnsemble(nsys, 16, 0, _native.MODE_FAST)
ens.set_params(e["dt"], e["eps"], e["G"])
ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
ens.step(steps, fused=fused)
ens.step(steps, fused=fused)
ens.synchronize()
ens.close()
