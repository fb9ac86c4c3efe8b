set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "push_exchange and (col6 or col7 or col8)" 2>&1 | tail -5 > gpurun_out/g6_pytest.log
WAVE_SWEEP_PROFILE=1 timeout 600 python tools/wave_sweep.py --variants "env:SPCSC_COL3=4;env:SPCSC_COL3=6;env:SPCSC_COL3=7;env:SPCSC_COL3=8" > gpurun_out/g6_wave.log 2>&1
cat gpurun_out/g6_pytest.log
