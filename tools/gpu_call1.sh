#!/bin/bash
# first GPU call of round 2: the CTA-pair syrk -- correctness, then timing against the single-CTA kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "syrk" > gpurun_out/r2_syrk_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_syrk_tests.log
timeout 300 python tools/perf_probe.py syrk2 > gpurun_out/r2_syrk2_perf.log 2>&1
echo "probe exit $?" >> gpurun_out/r2_syrk2_perf.log
