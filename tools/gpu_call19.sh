#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -k "hinv or chain or headline" > gpurun_out/r2_tests19.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests19.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_la.json 2> gpurun_out/r2_bench_la.err
echo "bench exit $?" >> gpurun_out/r2_bench_la.err
GPTQ_B200_CHAIN_LOOKAHEAD=0 timeout 400 python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_rec.json 2> gpurun_out/r2_bench_rec.err
echo "bench exit $?" >> gpurun_out/r2_bench_rec.err
timeout 400 python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_la2.json 2> gpurun_out/r2_bench_la2.err
