set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py -x -q -k "joint or cfg3 or clr or fresh or col6" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py cfg3a cfg3b 2>&1 | tail -2
