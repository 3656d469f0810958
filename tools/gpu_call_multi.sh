#!/bin/bash
# usage: tools/gpu_call_multi.sh N   (run under gpurun --gpus N)
N=$1
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench exit $?" >> gpurun_out/r2_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --workload model > gpurun_out/r2_model_n$N.json 2> gpurun_out/r2_model_n$N.err
echo "model exit $?" >> gpurun_out/r2_model_n$N.err
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2_tests_multi.log 2>&1
  echo "pytest exit $?" >> gpurun_out/r2_tests_multi.log
fi
