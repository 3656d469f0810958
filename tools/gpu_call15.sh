#!/bin/bash
set -x
timeout 300 python tools/chain_lab.py 4096 11008 > gpurun_out/r2_chain_lab.jsonl 2> gpurun_out/r2_chain_lab.err
echo "lab exit $?" >> gpurun_out/r2_chain_lab.err
for skip in 1 2 3; do
GPTQ_B200_CHAIN_SKIP=$skip timeout 300 python tools/chain_lab.py 4096 11008 >> gpurun_out/r2_chain_lab.jsonl 2>> gpurun_out/r2_chain_lab.err
done
GPTQ_B200_PDL=0 timeout 300 python tools/chain_lab.py 4096 11008 > gpurun_out/r2_chain_lab_nopdl.jsonl 2>> gpurun_out/r2_chain_lab.err
