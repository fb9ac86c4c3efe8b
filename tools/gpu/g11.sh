set -x
bash tools/ncu_capture.sh r02c 'k_col5|k_row_inv_prox3'
python tools/ncu_summary.py gpurun_out/r02c_raw.csv r02c "tools/ncu_capture.sh r02c 'k_col5|k_row_inv_prox3'" gpurun_out/r02c_source.csv > gpurun_out/r02c_summary_stdout.log 2>&1
cp profiles/r02c_ncu_summary.md profiles/r02_traffic.json gpurun_out/ 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/r02c_launches_bench.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driver.json 2> gpurun_out/r02_bench_driver.err

timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu --no-configs > gpurun_out/r02_bench_long.json 2> gpurun_out/r02_bench_long.err
ls -la gpurun_out
