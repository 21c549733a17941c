set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sample_split.py -q -m gpu 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check.log 2>&1; tail -25 gpurun_out/r2_multi_gpu_check.log
