set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "cdl_ms or cdl_bt or backtracking_golden or consensus or dictionary_learning" 2>&1 | tail -4
