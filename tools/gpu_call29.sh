#!/bin/bash
timeout 300 python tools/chain_sweep.py 16384 8 16 32 > gpurun_out/r2_chain_sweep4.jsonl 2> gpurun_out/r2_chain_sweep4.err
timeout 400 python tools/chain_sweep.py 36864 8 16 32 64 >> gpurun_out/r2_chain_sweep4.jsonl 2>> gpurun_out/r2_chain_sweep4.err
