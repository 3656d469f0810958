#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -k "hinv or chain" > gpurun_out/r2_tests26.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests26.log
timeout 300 python tools/chain_sweep.py 4096 1 2 > gpurun_out/r2_chain_sweep3.jsonl 2> gpurun_out/r2_chain_sweep3.err
timeout 300 python tools/chain_sweep.py 11008 4 6 8 12 >> gpurun_out/r2_chain_sweep3.jsonl 2>> gpurun_out/r2_chain_sweep3.err
timeout 300 python tools/chain_sweep.py 8192 2 4 8 >> gpurun_out/r2_chain_sweep3.jsonl 2>> gpurun_out/r2_chain_sweep3.err
timeout 300 python tools/chain_sweep.py 5120 1 2 4 >> gpurun_out/r2_chain_sweep3.jsonl 2>> gpurun_out/r2_chain_sweep3.err
