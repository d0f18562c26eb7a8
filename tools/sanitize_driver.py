#!/usr/bin/env python
"""Small workload touching every kernel family once (for compute-sanitizer runs; see profiles/)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "orbital-physics_b200")); sys.path.insert(0, REPO)
from core import _native, synthetic
from core.ensemble import EnsembleEngine

G = 6.67430e-11
def system(n, mode, radius=0.0, contacts=False, seed=1):
    c = synthetic.random_cloud(n, seed=seed, radius=radius)
    dev = _native.DeviceSystem(n, 0, mode)
    dev.set_params(c["dt"], c["eps"], G)
    dev.set_contacts(0.9, contacts)
    dev.set_history(4)
    dev.upload(*c.arrays())
    dev.accel(); dev.history_append()
    dev.step(3)
    dev.potential(); dev.energy_angmom(); dev.download_state(); dev.history_download(4)
    dev.close()

for n in (15, 100, 700):                       # micro / tiny / multi-kernel faithful paths, with contacts
    system(n, _native.MODE_FAITHFUL, radius=2e9, contacts=True, seed=n)
for n in (300, 3000, 9001):                    # symmetric fast kernel (TI 1, tails), detection on/off
    system(n, _native.MODE_FAST, radius=0.0, seed=n)
    system(n, _native.MODE_FAST, radius=2e8, contacts=True, seed=n + 1)
os.environ["ORBITAL_B200_SYM"] = "0"
system(3000, _native.MODE_FAST, radius=2e8, seed=5)          # one-sided fast kernel
os.environ["ORBITAL_B200_SYM"] = "1"
os.environ["ORBITAL_B200_SYM_TI"] = "8"
system(5000, _native.MODE_FAST, seed=6)
e = synthetic.ensemble(64, 16)
for mode in ("fast", "faithful"):
    ens = EnsembleEngine(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")), dt=e["dt"], softening=e["eps"], mode=mode)
    ens.step(3, fused=True); ens.step(2, fused=False); ens.energy(); ens.state(); ens.close()
print("sanitize driver ok")
