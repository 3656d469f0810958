#!/bin/bash
N=$1
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$i bench.py --gpus $N --steps 8 --warmup 3 --no-e2e > gpurun_out/r2_w3_n${N}_$i.json 2> gpurun_out/r2_w3_n${N}_$i.err
echo "exit $?" >> gpurun_out/r2_w3_n${N}_$i.err
done
