set -x
SPCSC_COL3=4 SPCSC_COL4_DBG=3 timeout 300 python tools/wave_sweep.py --variants "env:SPCSC_COL3=4,SPCSC_COL4_DBG=3" > gpurun_out/g5_dbg.log 2>&1
WAVE_SWEEP_PROFILE=1 timeout 600 python tools/wave_sweep.py --variants "env:SPCSC_COL3=4;env:SPCSC_COL3=6;f:2,1;f:4,1;f:4,2" > gpurun_out/g5_wave.log 2>&1
grep phases gpurun_out/g5_dbg.log | head
