#!/bin/bash
# does a longer replayed graph amortise a fixed per-graph-launch cost?  (un-fused ensemble, small-N fast steps)
for g in 16 64 256; do
  echo "== ensemble, graph of $g steps"
  ORBITAL_B200_ENS_GRAPH_STEPS=$g python - <<'PY'
import os, sys
sys.path.insert(0, "orbital-physics_b200")
import torch
from core import _native, synthetic
steps = 1024
for nsys in (4096, 8192, 16384, 65536):
    e = synthetic.ensemble_fast(nsys, 16)
    ens = _native.DeviceEnsemble(nsys, 16, 0, _native.MODE_FAST)
    ens.set_stream(torch.cuda.current_stream().cuda_stream)
    ens.set_params(e["dt"], e["eps"], e["G"])
    ens.upload(*(e[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m")))
    ens.step(steps, fused=False); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ens.step(steps, fused=False); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(nsys, round(1e3 * best / steps, 3), "us/step")
    ens.close()
PY
done
for g in 16 64 256; do
  echo "== fast N-body step, graph of $g steps"
  ORBITAL_B200_GRAPH_STEPS=$g python tools/sweep_step.py uniform 1024 4096 2>&1 | grep default
done
