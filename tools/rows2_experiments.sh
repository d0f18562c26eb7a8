#!/bin/bash
# timing experiment on the pass-2 producer/consumer kernel: mbarrier try_wait (may suspend) vs a test_wait spin
for v in "" vS; do
  lib=orbital-physics_b200/csrc/liborbital_b200${v:+_$v}.so
  echo "== ${v:-default} $lib"
  ORBITAL_B200_LIB=$PWD/$lib python tools/sweep_faithful_step.py 2>&1 | tail -9
done
ORBITAL_B200_LIB=$PWD/orbital-physics_b200/csrc/liborbital_b200_vS.so python -m pytest tests/test_device.py -x -q -m gpu -k "faithful or disk" 2>&1 | tail -2
