#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_api_edges.py tests/test_gpu_e2e_model.py tests/test_gpu_checkpoint.py -q -s --maxfail=20 > gpurun_out/r2_tests5.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests5.log
GPTQ_B200_CHUNK_TOKENS=1024 timeout 600 python -m pytest tests/test_gpu_headline.py -q -s -k "llama7b or opt66b or bloom" > gpurun_out/r2_tests5_chunk1024.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests5_chunk1024.log
timeout 300 python tools/gap_profile.py > gpurun_out/r2_gaps.json 2> gpurun_out/r2_gaps.err
timeout 300 python tools/perf_probe.py sgemm hinv hinvbig > gpurun_out/r2_probe5.log 2>&1
