set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "push_exchange or staged_column" 2>&1 | tail -5 > gpurun_out/g2_pytest.log
WAVE_SWEEP_PROFILE=1 timeout 600 python tools/wave_sweep.py --variants "fused;env:SPCSC_COL3=4;env:SPCSC_COL3=11;env:SPCSC_COL3=12;env:SPCSC_COL3=3;env:SPCSC_COL3=2" > gpurun_out/g2_wave.log 2>&1
cat gpurun_out/g2_pytest.log
