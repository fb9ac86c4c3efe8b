set -x
mkdir -p /tmp/ncu
for cfg in cfg3a cfg4 cfg5 cfg5_cns; do
  ncu --set full --clock-control none -k regex:"k_row_inv_prox3|k_row_inv_prox2|k_col2|k_col5|k_row_prox_fwd3|k_ccmod_grad|k_pcn|k_cns_|k_row_inv|k_row_fwd|k_pgm_momentum|k_spec_diffnorm|k_ccmod_step" -s 60 -c 24 -f -o /tmp/ncu/$cfg python tools/bench_configs.py $cfg > gpurun_out/g15_$cfg.log 2>&1
  ncu -i /tmp/ncu/$cfg.ncu-rep --page raw --csv > gpurun_out/g15_${cfg}_raw.csv 2>> gpurun_out/g15_$cfg.log
done
python tools/ncu_configs_summary.py gpurun_out/r02_configs_ncu.md cfg3a=gpurun_out/g15_cfg3a_raw.csv cfg4=gpurun_out/g15_cfg4_raw.csv cfg5=gpurun_out/g15_cfg5_raw.csv cfg5_cns=gpurun_out/g15_cfg5_cns_raw.csv | tail -40
