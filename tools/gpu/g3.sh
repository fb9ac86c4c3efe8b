set -x
bash tools/ncu_capture.sh r02b 'k_col4' SPCSC_COL3=4
python tools/ncu_summary.py gpurun_out/r02b_raw.csv r02b_box "tools/ncu_capture.sh r02b k_col4 SPCSC_COL3=4" gpurun_out/r02b_source.csv > gpurun_out/r02b_summary_stdout.log 2>&1
cp profiles/r02b_box_ncu_summary.md gpurun_out/ 2>/dev/null
ls -la gpurun_out
