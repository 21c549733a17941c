# Round 1 (k): launch list of the from_activations leg of bench.py, VGG16 full pass (one GPU, and one rank of eight).
# Run under gpurun.  (profiles/r1k_conv_corr9_vgg_conv1.md came from:
#   ncu --set full --clock-control none --import-source on -k regex:conv_corr9_tma_kernel -c 2 -o gpurun_out/r1k_corr9_vgg_conv1 \
#       python tools/vgg_bench.py --skip-dense --n-img 376 --reps 1 --layers 1)
set -x
K='regex:conv_|corr9|gemm_nt|sweep_|dense_stream|transpose_|reduce_splits|gram_i8|i8_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2000 --csv --log-file gpurun_out/r1k_launches_activations.csv python bench.py --steps 2 --warmup 3 --profile-activations-leg > gpurun_out/r1k_launches_activations.log 2>&1
python tools/vgg_bench.py --reps 2 > gpurun_out/r1k_vgg_full.log 2>&1
python tools/vgg_bench.py --reps 2 --shard 3/8 > gpurun_out/r1k_vgg_shard3of8.log 2>&1
