for cfg in cfg5_cns cfg5; do
timeout 120 python bench.py --config $cfg --steps 40 --warmup 10 > gpurun_out/r02_bench_${cfg}_n1.json 2> gpurun_out/r02_bench_${cfg}_n1.err
cut -c1-260 gpurun_out/r02_bench_${cfg}_n1.json
done
