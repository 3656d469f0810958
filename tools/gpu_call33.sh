#!/bin/bash
N=8
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
cat /sys/devices/system/node/online > gpurun_out/r2_numa_nodes_n$N.txt 2>&1
python -c "import os;print(sorted(os.sched_getaffinity(0)))" >> gpurun_out/r2_numa_nodes_n$N.txt 2>&1
GPTQ_B200_NUMA_BIND=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_numa_n${N}_b1.json 2> gpurun_out/r2_numa_n${N}_b1.err
echo "exit $?" >> gpurun_out/r2_numa_n${N}_b1.err
