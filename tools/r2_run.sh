set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "lowrank or dense or golden or kat or shard or fc1 or fc2 or auto_picks" 2>&1 | tail -5
python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,25088x512x1504,2048x128x5008 --methods auto --reps 2 2>&1 | grep shape | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2l_bench_vgg_n4.json 2> gpurun_out/r2l_bench_vgg_n4.err; tail -c 1200 gpurun_out/r2l_bench_vgg_n4.json; tail -5 gpurun_out/r2l_bench_vgg_n4.err
