set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "consensus or cdl_cns" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py cfg5_cns 2>&1 | tail -2
