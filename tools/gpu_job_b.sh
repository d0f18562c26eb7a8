#!/bin/bash
# round-2 job B (N GPUs): NCCL parity test + the multi-GPU bench line
N=${1:-2}
python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_tests_nccl_g$N.log
cat gpurun_out/r2_tests_nccl_g$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err
tail -5 gpurun_out/r2_bench_g$N.err
python - $N <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2_bench_g{sys.argv[1]}.json"))
for k in ("value", "ms_per_step", "force_ms", "comm_ms", "parity_check", "gpu_launches", "strong_262144"):
    print(k, d.get(k))
print(d["e2e"], d["config"]["parallelism"])
print(d.get("ensemble"))
PY
