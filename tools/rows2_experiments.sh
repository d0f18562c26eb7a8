#!/bin/bash
# timing experiments on the pass-2 producer/consumer kernel (variants built with -D flags; B and C give wrong results)
for v in "" vA vB vC vD vE; do
  lib=orbital-physics_b200/csrc/liborbital_b200${v:+_$v}.so
  echo "== ${v:-default} $lib"
  ORBITAL_B200_LIB=$PWD/$lib python - <<'PY'
import os, sys, time
sys.path.insert(0, "orbital-physics_b200")
from core import _native, synthetic
for n in (2048, 4096):
    c = synthetic.random_cloud(n, seed=n)
    dev = _native.DeviceSystem(n, 0, _native.MODE_FAITHFUL)
    dev.set_params(c["dt"], c["eps"], c["G"]); dev.upload(*c.arrays()); dev.accel(); dev.step(16)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); dev.step(64); best = min(best, time.perf_counter() - t0)
    print(n, round(1e6 * best / 64, 2), "us/step")
    dev.close()
PY
done
