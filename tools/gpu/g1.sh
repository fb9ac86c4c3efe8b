set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/g1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g1_bench_driver.json 2> gpurun_out/g1_bench_driver.err
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu --no-configs > gpurun_out/g1_bench_long.json 2> gpurun_out/g1_bench_long.err
WAVE_SWEEP_PROFILE=1 timeout 300 python tools/wave_sweep.py --variants "fused;2,2;3,2;4,2" > gpurun_out/g1_wave.log 2>&1
SPCSC_COL3=1 timeout 200 python bench.py --steps 200 --warmup 50 --no-cpu --no-configs > gpurun_out/g1_bench_col3.json 2>&1
tail -3 gpurun_out/g1_pytest.log
