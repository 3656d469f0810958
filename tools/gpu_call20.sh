#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests20.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests20.log
timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_la3.json 2> gpurun_out/r2_bench_la3.err
echo "bench exit $?" >> gpurun_out/r2_bench_la3.err
GPTQ_B200_EXPANDABLE_SEGMENTS=0 timeout 400 python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_la4.json 2> gpurun_out/r2_bench_la4.err
echo "bench exit $?" >> gpurun_out/r2_bench_la4.err
