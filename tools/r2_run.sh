set -x
mkdir -p gpurun_out
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_vgg16_n$N.json 2> gpurun_out/r2_bench_vgg16_n$N.err; tail -3 gpurun_out/r2_bench_vgg16_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cifar10_cnn --steps 10 --warmup 3 > gpurun_out/r2_bench_cifar10_cnn_n$N.json 2> /dev/null
