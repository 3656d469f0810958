#!/bin/bash
for i in 1 2 3; do
timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline >> gpurun_out/r2_bench_sampler3.jsonl 2>> gpurun_out/r2_bench_sampler3.err
done
GPTQ_B200_CLOCK_SAMPLER=smi timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline >> gpurun_out/r2_bench_sampler3.jsonl 2>> gpurun_out/r2_bench_sampler3.err
