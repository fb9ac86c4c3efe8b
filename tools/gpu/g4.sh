set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "push_exchange and col5" 2>&1 | tail -5 > gpurun_out/g4_pytest.log
WAVE_SWEEP_PROFILE=1 timeout 600 python tools/wave_sweep.py --variants "env:SPCSC_COL3=4;env:SPCSC_COL3=5;env:SPCSC_COL3=5,SPCSC_COL5_STAGGER=1500;env:SPCSC_COL3=5,SPCSC_COL5_STAGGER=3000" > gpurun_out/g4_wave.log 2>&1
cat gpurun_out/g4_pytest.log
