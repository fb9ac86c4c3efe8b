set -x
N=2
for v in 6 0; do
SPCSC_COL3=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 50 > gpurun_out/g14_n2_col$v.json 2> gpurun_out/g14_n2_col$v.err
done
timeout 300 python bench.py --steps 200 --warmup 50 --no-cpu --no-configs > gpurun_out/g14_n1.json 2>&1
