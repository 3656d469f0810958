#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -k "dequant or tril or two_gpu" > gpurun_out/r2_tests12.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests12.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e > gpurun_out/r2_bench12_n2.json 2> gpurun_out/r2_bench12_n2.err
echo "bench exit $?" >> gpurun_out/r2_bench12_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload model --model-layers 6 > gpurun_out/r2_model12_n2.json 2> gpurun_out/r2_model12_n2.err
echo "model exit $?" >> gpurun_out/r2_model12_n2.err
