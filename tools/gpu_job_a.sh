#!/bin/bash
# round-2 job A: sharded + engine GPU tests, then the default bench line
python -m pytest tests/test_sharded_gpu.py tests/test_engine.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_tests_a.log
cat gpurun_out/r2_tests_a.log
python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -5 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_a.json"))
print({k: d[k] for k in ("value", "ms_per_step", "parity_check", "gpu_launches")})
print(d["roofline"]["frac"], d["e2e"]["value"])
print(json.dumps(d["scale_denominator"]))
print(json.dumps(d["configs"], indent=1))
print(json.dumps(d["cpu_baseline"])[:1500])
print(d["ensemble"])
PY
