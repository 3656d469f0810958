#!/bin/bash
# the 8-GPU box: strong scaling of the layer bench at N = 1, 2, 4, 8 back to back, then the whole model on 8 GPUs
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_gpus_n8.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
echo "n1 exit $?" >> gpurun_out/r2_scale_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  echo "n$N exit $?" >> gpurun_out/r2_scale_n$N.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --workload model > gpurun_out/r2_model_n8.json 2> gpurun_out/r2_model_n8.err
echo "model exit $?" >> gpurun_out/r2_model_n8.err
