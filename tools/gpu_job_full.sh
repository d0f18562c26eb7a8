#!/bin/bash
# full single-GPU verification: all GPU tests, smoke, fuzz, the default bench line, the reference arm
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_gputests_full.log; cat gpurun_out/r2_gputests_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/fuzz_faithful.py 45 11 2>&1 | tail -1 | tee gpurun_out/r2_fuzz.txt
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -3 gpurun_out/r2_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> /dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "parity_check", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"])
print([ (r["n_bodies"], r["velocity_dtype"], round(r["us_per_step"],3), r["bit_exact_vs_oracle_after_10100_steps"]) for r in d["configs"]["C0_solar_system"]["runs"]])
print(d["configs"]["C0_solar_system"]["live_reference_check"])
print(d["configs"]["C1_disk_4096"]["runs"])
print(d["ensemble"])
r = json.load(open("gpurun_out/r2_bench_reference_arm.json"))
print(d["ic_pipeline"])
print("step() loop us/step:", [r.get("step_call_loop_us_per_step") for r in d["configs"]["C0_solar_system"]["runs"]])
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
