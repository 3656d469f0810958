#!/bin/bash
GPTQ_B200_GEMM_CPREFETCH=0 timeout 300 python tools/gemm_lab.py > gpurun_out/r2_gemm_lab.jsonl 2> gpurun_out/r2_gemm_lab.err
GPTQ_B200_GEMM_CPREFETCH=1 timeout 300 python tools/gemm_lab.py >> gpurun_out/r2_gemm_lab.jsonl 2>> gpurun_out/r2_gemm_lab.err
GPTQ_B200_GEMM_CPREFETCH=0 timeout 300 python tools/gemm_lab.py >> gpurun_out/r2_gemm_lab.jsonl 2>> gpurun_out/r2_gemm_lab.err
GPTQ_B200_GEMM_CPREFETCH=1 timeout 300 python tools/gemm_lab.py >> gpurun_out/r2_gemm_lab.jsonl 2>> gpurun_out/r2_gemm_lab.err
