#!/bin/bash
for i in 1 2 3 4; do
timeout 400 python bench.py --steps 25 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline >> gpurun_out/r2_bench_sampler4.jsonl 2>> gpurun_out/r2_bench_sampler4.err
done
