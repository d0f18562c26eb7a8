#!/bin/bash
# Development aid: ensemble throughput vs systems per CTA.
cd "$(dirname "$0")/.."
for w in 1 2 4 8; do
ORBITAL_B200_ENS_WARPS=$w python - <<PY
import sys; sys.path.insert(0, "orbital-physics_b200"); sys.path.insert(0, ".")
import torch, bench
r = bench.ensemble_measure(0, torch, steps=200)
print("warps/CTA=$w  unfused %.0f GB/s  (%.3e system-steps/s)  fused %.3e int/s" % (r["unfused_gbs"], r["unfused_system_steps_per_s"], r["fused_interactions_per_s"]))
PY
done
