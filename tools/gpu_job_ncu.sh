#!/bin/bash
# round-2 ncu evidence (one GPU): launch list of the bench command, then --set full captures of the hot kernels
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r2_bench_for_ncu.json 2> gpurun_out/r2_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/ncu_a.log 2>&1
python tools/ens_driver.py 65536 fused > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ens_step_fast -s 2 -c 1 -o gpurun_out/r2_ens_fused python tools/ens_driver.py 65536 fused > gpurun_out/ncu_b.log 2>&1
python tools/ens_driver.py 524288 unfused > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ens_step_fast -s 20 -c 1 -o gpurun_out/r2_ens_unfused_hbm python tools/ens_driver.py 524288 unfused > gpurun_out/ncu_c.log 2>&1
python tools/ens_driver.py 8192 fused 256 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ens_fast_sliced -s 1 -c 1 -o gpurun_out/r2_ens_sliced python tools/ens_driver.py 8192 fused 256 > gpurun_out/ncu_d.log 2>&1
python tools/step_driver.py 262144 fast > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:force_sym -s 2 -c 1 -o gpurun_out/r2_force_sym python tools/step_driver.py 262144 fast > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:reduce_sym -s 2 -c 1 -o gpurun_out/r2_reduce_sym python tools/step_driver.py 4096 fast > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:faithful_pairs -s 1 -c 1 -o gpurun_out/r2_faithful_pairs python tools/step_driver.py 4096 > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:faithful_rows2 -s 1 -c 1 -o gpurun_out/r2_faithful_rows2 python tools/step_driver.py 4096 > gpurun_out/ncu_h.log 2>&1
ls -la gpurun_out/*.ncu-rep
