set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/g10_pytest.log
cat gpurun_out/g10_pytest.log
timeout 300 python tools/bench_configs.py cfg5 cfg5_cns > gpurun_out/g10_cfg5.log 2>&1
cat gpurun_out/g10_cfg5.log
