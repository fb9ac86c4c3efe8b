set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/g23_pytest.log
cat gpurun_out/g23_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs 2>/dev/null | cut -c1-400
