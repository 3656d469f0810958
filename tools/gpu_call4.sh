#!/bin/bash
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=20 > gpurun_out/r2_tests_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests_full.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench_n1.err
