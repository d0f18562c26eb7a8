#!/usr/bin/env python
"""Tiny driver for profiling the ensemble kernel: 65,536 x 16, a few un-fused steps."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
from core import _native, synthetic
e = synthetic.ensemble_fast(65536, 16)
ens = _native.DeviceEnsemble(65536, 16, 0, _native.MODE_FAST)
ens.set_params(e["dt"], e["eps"], e["G"])
ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
ens.step(6, fused=False)
ens.synchronize()
print("ok", ens.launch_count())
