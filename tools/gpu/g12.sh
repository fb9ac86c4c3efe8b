set -x
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 1500 gpurun_out/r02_bench_n$N.json
tail -5 gpurun_out/r02_bench_n$N.err
