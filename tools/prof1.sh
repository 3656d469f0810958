set -x
mkdir -p gpurun_out
python tools/perf_probe.py hinv colloop sgemm > gpurun_out/probe_base.log 2>&1
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"gemm_tc|chol_diag" -c 1500 --csv --log-file gpurun_out/hinv_launches.csv python tools/perf_probe.py hinv > gpurun_out/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blockquant -s 40 -c 2 -f -o gpurun_out/prof_bq python tools/perf_probe.py colloop > gpurun_out/ncu_bq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_diag -s 5 -c 1 -f -o gpurun_out/prof_diag python tools/perf_probe.py hinv > gpurun_out/ncu_diag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_gemm python tools/perf_probe.py sgemm > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
