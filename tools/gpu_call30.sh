#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "gemm or hinv or lazy or colloop" > gpurun_out/r2_tests30.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests30.log
GPTQ_B200_GEMM_COALESCED=0 timeout 300 python tools/gemm_lab.py > gpurun_out/r2_gemm_lab2.jsonl 2> gpurun_out/r2_gemm_lab2.err
GPTQ_B200_GEMM_COALESCED=1 timeout 300 python tools/gemm_lab.py >> gpurun_out/r2_gemm_lab2.jsonl 2>> gpurun_out/r2_gemm_lab2.err
timeout 300 python tools/chain_sweep.py 4096 0 > gpurun_out/r2_chain_sweep5.jsonl 2>> gpurun_out/r2_gemm_lab2.err
timeout 300 python tools/chain_sweep.py 11008 0 >> gpurun_out/r2_chain_sweep5.jsonl 2>> gpurun_out/r2_gemm_lab2.err
timeout 400 python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_co.json 2> gpurun_out/r2_bench_co.err
GPTQ_B200_GEMM_COALESCED=0 timeout 400 python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_noco.json 2> gpurun_out/r2_bench_noco.err
