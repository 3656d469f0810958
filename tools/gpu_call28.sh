#!/bin/bash
# final 1-GPU pass of round 2: full GPU test suite, the bench with every arm, the other configs, the whole model, ncu
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_tests_final.log
timeout 900 python bench.py --steps 10 --warmup 5 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench_final_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref exit $?" >> gpurun_out/r2_bench_ref.err
for cfg in opt125m opt1.3b bloom7b1; do
timeout 600 python bench.py --config $cfg --steps 6 --warmup 3 --no-gpu-baseline > gpurun_out/r2_bench_$cfg.json 2> gpurun_out/r2_bench_$cfg.err
echo "bench $cfg exit $?" >> gpurun_out/r2_bench_$cfg.err
done
timeout 600 python bench.py --config opt66b-layer --steps 3 --warmup 2 --no-gpu-baseline --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_opt66b.json 2> gpurun_out/r2_bench_opt66b.err
echo "bench opt66b exit $?" >> gpurun_out/r2_bench_opt66b.err
timeout 600 python bench.py --workload model > gpurun_out/r2_model_n1.json 2> gpurun_out/r2_model_n1.err
echo "model exit $?" >> gpurun_out/r2_model_n1.err
timeout 300 python tools/perf_probe.py hbm once > gpurun_out/r2_hbm_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:"pack_kernel|pack_weights|find_params_kernel|vecquant_kernel" -c 18 -o gpurun_out/r2_prof_hbm2 python tools/perf_probe.py hbm once > gpurun_out/r2_ncu_hbm2.log 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_plain_for_ncu.log 2>&1 && \
timeout 1000 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 3400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2_ncu_list.log
ls -la gpurun_out/ | head -50
