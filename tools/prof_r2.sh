#!/bin/bash
# Round-2 profiling evidence (run under gpurun, one GPU).  Outputs under gpurun_out/, summarised into profiles/ by
# tools/launch_summary.py / tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
# 0. bench lines of the three single-GPU workloads (the numbers; nothing below is a bench value)
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_vgg16_n1.json 2> gpurun_out/r2_bench_vgg16_n1.err
python bench.py --workload cifar10_cnn --steps 10 --warmup 3 > gpurun_out/r2_bench_cifar10_cnn_n1.json 2> /dev/null
python bench.py --workload mnist_mlp --steps 10 --warmup 3 > gpurun_out/r2_bench_mnist_mlp_n1.json 2> /dev/null
# 1. launch list of the VGG16 value leg (3 warm-up passes + 1 timed), full size
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2_bench_launches.log 2>&1
# 2. the dominant kernel (correlation-form conv Grams), ncu --set full, 376 images (ncu saves / restores device memory per replay)
ncu --set full --clock-control none --import-source on -k regex:conv_corr9_tma -s 2 -c 4 -o gpurun_out/r2_corr9_vgg -f \
    python bench.py --profile --steps 1 --warmup 3 --n-img 376 > gpurun_out/r2_corr9_vgg.log 2>&1
# 2b. the strip kernel of the small images on VGG16's 28 x 28 x 512 layer (the xq x x pass of conv8)
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:conv_corr9_strip_kernel -s 33 -c 2 -o gpurun_out/r2_strip_conv8 -f \
    python bench.py --profile --steps 1 --warmup 1 --n-img 376 > gpurun_out/r2_strip_conv8.log 2>&1
# 3. the tcgen05 contraction and the tensor-core range walk of the residual-form sweep on VGG16 fc1
ncu --set full --clock-control none --import-source on -k regex:slgemm_i8 -s 60 -c 3 -o gpurun_out/r2_slgemm_fc1 -f \
    python tools/dense_bench.py --shapes 25088x4096x1504 --methods auto --reps 0 > gpurun_out/r2_slgemm_fc1.log 2>&1
ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:sweep_tc_kernel -s 20 -c 1 -o gpurun_out/r2_tc_fc1 -f \
    python tools/dense_bench.py --shapes 25088x4096x1504 --methods auto --reps 0 > gpurun_out/r2_tc_fc1.log 2>&1
# 4. every Dense method on the shapes of configs 1-4; the FP64 / LDS dispatch microbenchmark behind the walk's cost model
python tools/dense_methods.py --out gpurun_out/dense_methods_r2.md --reps 3 > gpurun_out/dense_methods_r2.log 2>&1
./tools/bin/fp64_issue > gpurun_out/fp64_issue_r2.txt 2>&1
ls -la gpurun_out/*.ncu-rep
