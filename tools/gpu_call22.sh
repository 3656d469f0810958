#!/bin/bash
for kind in nvml smi off nvml; do
GPTQ_B200_CLOCK_SAMPLER=$kind timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline >> gpurun_out/r2_bench_sampler.jsonl 2>> gpurun_out/r2_bench_sampler.err
done
