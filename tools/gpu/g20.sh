set -x
N=${NGPU:-2}
for cfg in cfg5_cns cfg5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --config $cfg --gpus $N --steps 40 --warmup 10 > gpurun_out/r02_bench_${cfg}_n$N.json 2> gpurun_out/r02_bench_${cfg}_n$N.err
tail -c 900 gpurun_out/r02_bench_${cfg}_n$N.json; grep -i "error\|Traceback" -A 5 gpurun_out/r02_bench_${cfg}_n$N.err | tail -12
done
