#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -k "hinv or chain" > gpurun_out/r2_tests14.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests14.log
timeout 400 python tools/chain_lab.py 4096 11008 > gpurun_out/r2_chain_lab.jsonl 2> gpurun_out/r2_chain_lab.err
echo "lab exit $?" >> gpurun_out/r2_chain_lab.err
timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_la.json 2> gpurun_out/r2_bench_la.err
echo "bench exit $?" >> gpurun_out/r2_bench_la.err
GPTQ_B200_CHAIN_LOOKAHEAD=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_rec.json 2> gpurun_out/r2_bench_rec.err
echo "bench exit $?" >> gpurun_out/r2_bench_rec.err
