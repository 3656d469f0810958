# Round profile capture (run under gpurun, one GPU): bench lines first (no profiler), then the ncu launch list of one
# bench step and ncu --set full captures of the dominant kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1580 -c 4000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_tc -s 8 -c 1 -f -o gpurun_out/prof_syrk_k4096 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_syrk1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:syrk_tc -s 28 -c 1 -f -o gpurun_out/prof_syrk_k11008 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_syrk2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm python tools/perf_probe.py sgemm > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blockquant -s 40 -c 1 -f -o gpurun_out/prof_bq python tools/perf_probe.py colloop > gpurun_out/ncu_bq.log 2>&1
ls -la gpurun_out | tail -20
