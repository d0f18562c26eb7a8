#!/usr/bin/env python
"""Tiny driver for profiling the ensemble kernel: NSYS x 16 (default 65,536), a few un-fused steps.

    python tools/ens_profile.py [nsys]      # 524288 -> 671 MB of state, 5x the L2: the HBM-bound regime
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbital-physics_b200"))
from core import _native, synthetic
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
e = synthetic.ensemble_fast(nsys, 16)
ens = _native.DeviceEnsemble(nsys, 16, 0, _native.MODE_FAST)
ens.set_params(e["dt"], e["eps"], e["G"])
ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
ens.step(6, fused=False)
ens.synchronize()
print("ok", ens.launch_count())
