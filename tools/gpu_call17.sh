#!/bin/bash
timeout 300 python tools/chain_sweep.py 11008 6 8 12 16 > gpurun_out/r2_chain_sweep.jsonl 2> gpurun_out/r2_chain_sweep.err
GPTQ_B200_CHAIN_SKIP=3 timeout 300 python tools/chain_sweep.py 11008 8 16 >> gpurun_out/r2_chain_sweep.jsonl 2>> gpurun_out/r2_chain_sweep.err
GPTQ_B200_CHAIN_TRACE=1 timeout 300 python tools/chain_trace.py 11008 8 > gpurun_out/r2_chain_trace_8.out 2> gpurun_out/r2_chain_trace_8.err
