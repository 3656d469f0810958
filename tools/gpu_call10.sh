#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_edges.py -q -k "quantlinear or pack or tril or api_edges" > gpurun_out/r2_tests10.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests10.log
timeout 900 python bench.py --workload model --config llama7b > gpurun_out/r2_model_n1.json 2> gpurun_out/r2_model_n1.err
echo "model exit $?" >> gpurun_out/r2_model_n1.err
for cfg in opt125m opt1.3b; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-gpu-baseline > gpurun_out/r2_bench_$cfg.json 2> gpurun_out/r2_bench_$cfg.err
  echo "bench $cfg exit $?" >> gpurun_out/r2_bench_$cfg.err
done
timeout 600 python bench.py --config bloom7b1 --steps 3 --warmup 2 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2_bench_bloom7b1.json 2> gpurun_out/r2_bench_bloom7b1.err
echo "bench bloom exit $?" >> gpurun_out/r2_bench_bloom7b1.err
timeout 900 python bench.py --config opt66b-layer --steps 2 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_opt66b.json 2> gpurun_out/r2_bench_opt66b.err
echo "bench opt66b exit $?" >> gpurun_out/r2_bench_opt66b.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench10_n1.json 2> gpurun_out/r2_bench10_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench10_n1.err
