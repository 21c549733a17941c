set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2m_bench_vgg_n8.json 2> gpurun_out/r2m_bench_vgg_n8.err; tail -c 800 gpurun_out/r2m_bench_vgg_n8.json; tail -5 gpurun_out/r2m_bench_vgg_n8.err
