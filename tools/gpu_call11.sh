#!/bin/bash
set -x
timeout 600 python bench.py --workload model --config llama7b --model-layers 6 > gpurun_out/r2_model6_base.json 2> gpurun_out/r2_model6_base.err
GPTQ_B200_EXPANDABLE_SEGMENTS=1 timeout 600 python bench.py --workload model --config llama7b --model-layers 6 > gpurun_out/r2_model6_exp.json 2> gpurun_out/r2_model6_exp.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench11_n1.json 2> gpurun_out/r2_bench11_n1.err
