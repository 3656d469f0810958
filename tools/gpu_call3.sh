#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "syrk" > gpurun_out/r2_syrk_tests3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_syrk_tests3.log
timeout 300 python tools/syrk_lab.py sustained 1.2 2,3 > gpurun_out/r2_syrk_lab3.log 2>&1
echo "lab exit $?" >> gpurun_out/r2_syrk_lab3.log
timeout 120 python tools/syrk_lab.py sweep > gpurun_out/r2_sweep_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:syrk_tc2 --csv --log-file gpurun_out/r2_sweep_ncu.csv python tools/syrk_lab.py sweep > gpurun_out/r2_sweep_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_sweep_ncu.log
