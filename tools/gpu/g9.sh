set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/g9_multi.log 2>&1
grep -v "^\[W\|Setting OMP\|^\*\*\*" gpurun_out/g9_multi.log | head -30
