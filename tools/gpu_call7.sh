#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_api_edges.py -q > gpurun_out/r2_tests7.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests7.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench7_n1.json 2> gpurun_out/r2_bench7_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench7_n1.err
timeout 300 python tools/perf_probe.py hbm > gpurun_out/r2_probe_hbm.log 2>&1
timeout 120 python tools/syrk_lab.py once > gpurun_out/r2_once7.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_tc2 -c 10 -o gpurun_out/r2_prof_syrk_final python tools/syrk_lab.py once > gpurun_out/r2_ncu_syrk_final.log 2>&1
timeout 120 python tools/perf_probe.py hbm > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:"pack_kernel|pack_weights|find_params_kernel|vecquant_kernel" -s 30 -c 90 -o gpurun_out/r2_prof_hbm python tools/perf_probe.py hbm > gpurun_out/r2_ncu_hbm.log 2>&1
echo done
