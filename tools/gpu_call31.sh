#!/bin/bash
N=$1
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
for b in 1 0; do
GPTQ_B200_NUMA_BIND=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$b bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_numa_n${N}_b$b.json 2> gpurun_out/r2_numa_n${N}_b$b.err
echo "exit $?" >> gpurun_out/r2_numa_n${N}_b$b.err
done
