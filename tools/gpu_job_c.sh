#!/bin/bash
python -m pytest tests/test_device.py -x -q -m gpu -k "ensemble or legacy" 2>&1 | tail -12 > gpurun_out/r2_tests_c.log
cat gpurun_out/r2_tests_c.log
python tools/ens_sweep.py 16 > gpurun_out/r2_ens_sizes.txt 2>&1
cat gpurun_out/r2_ens_sizes.txt
