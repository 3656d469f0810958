#!/bin/bash
for K in 2048 5120 8192 9216; do
timeout 300 python tools/chain_sweep.py $K 1 2 4 8 >> gpurun_out/r2_chain_sweep2.jsonl 2>> gpurun_out/r2_chain_sweep2.err
done
GPTQ_B200_CHAIN_INV_PRIO=1 timeout 300 python tools/chain_sweep.py 11008 4 8 > gpurun_out/r2_chain_sweep_invprio.jsonl 2>> gpurun_out/r2_chain_sweep2.err
GPTQ_B200_CHAIN_INV_PRIO=1 timeout 300 python tools/chain_sweep.py 4096 1 2 >> gpurun_out/r2_chain_sweep_invprio.jsonl 2>> gpurun_out/r2_chain_sweep2.err
