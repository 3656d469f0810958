#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -k "syrk or headline_shape or add_batch" > gpurun_out/r2_tests13.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests13.log
timeout 900 python bench.py --config opt66b-layer --steps 2 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_opt66b.json 2> gpurun_out/r2_bench_opt66b.err
echo "bench opt66b exit $?" >> gpurun_out/r2_bench_opt66b.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench_final_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref exit $?" >> gpurun_out/r2_bench_ref.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_plain_for_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_ncu_list.log
ls -la gpurun_out/r2_launches.csv
