#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "syrk" > gpurun_out/r2_syrk_tests2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_syrk_tests2.log
timeout 300 python tools/syrk_lab.py sustained 1.5 1,2,3 > gpurun_out/r2_syrk_lab.log 2>&1
echo "lab exit $?" >> gpurun_out/r2_syrk_lab.log
timeout 120 python tools/syrk_lab.py once > gpurun_out/r2_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_tc -c 12 -o gpurun_out/r2_prof_syrk python tools/syrk_lab.py once > gpurun_out/r2_ncu_syrk.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_ncu_syrk.log
