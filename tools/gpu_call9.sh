#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_checkpoint.py -q > gpurun_out/r2_tests9.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests9.log
GPTQ_B200_CHUNK_TOKENS=1024 timeout 600 python -m pytest tests/test_gpu_headline.py -q -s -k "headline_shape" > gpurun_out/r2_tests9_chunk1024.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests9_chunk1024.log
timeout 300 python tools/syrk_lab.py l2 > gpurun_out/r2_syrk_l2.log 2>&1
timeout 300 python tools/perf_probe.py hbm > gpurun_out/r2_probe_hbm9.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench9_n1.json 2> gpurun_out/r2_bench9_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench9_n1.err
