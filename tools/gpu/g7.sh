set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/g7_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > gpurun_out/g7_bench_driver.json 2> gpurun_out/g7_bench_driver.err
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu --no-configs > gpurun_out/g7_bench_long.json 2> gpurun_out/g7_bench_long.err
cat gpurun_out/g7_pytest.log
