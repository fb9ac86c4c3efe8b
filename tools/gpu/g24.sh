set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "consensus or gradient_regularisation or backtracking_golden or dictionary_learning" 2>&1 | tail -3
