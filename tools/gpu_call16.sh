#!/bin/bash
for pb in 2 4; do
GPTQ_B200_CHAIN_TRACE=1 timeout 300 python tools/chain_trace.py 11008 $pb > gpurun_out/r2_chain_trace_$pb.out 2> gpurun_out/r2_chain_trace_$pb.err
done
GPTQ_B200_CHAIN_TRACE=1 timeout 300 python tools/chain_trace.py 4096 1 > gpurun_out/r2_chain_trace_4096_1.out 2> gpurun_out/r2_chain_trace_4096_1.err
