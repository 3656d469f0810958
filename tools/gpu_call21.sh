#!/bin/bash
timeout 300 python tools/stall_probe.py 25 > gpurun_out/r2_stall_la.log 2>&1
GPTQ_B200_CHAIN_LOOKAHEAD=0 timeout 300 python tools/stall_probe.py 25 > gpurun_out/r2_stall_rec.log 2>&1
GPTQ_B200_PDL=0 timeout 300 python tools/stall_probe.py 25 > gpurun_out/r2_stall_la_nopdl.log 2>&1
