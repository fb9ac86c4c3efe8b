set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "consensus or cdl_cns" 2>&1 | tail -5 > gpurun_out/g8_pytest.log
cat gpurun_out/g8_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/g8_multi.log 2>&1
tail -12 gpurun_out/g8_multi.log
