set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/g19_pytest.log
cat gpurun_out/g19_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g19_smoke.log 2>&1; tail -3 gpurun_out/g19_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g19_bench.json 2> gpurun_out/g19_bench.err; tail -c 300 gpurun_out/g19_bench.err
