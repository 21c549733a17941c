# Round 1 (k): profiles of the correlation-form conv Grams and the launch lists of both bench legs.  Run under gpurun.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1k_launches_value.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-activations-leg > gpurun_out/r1k_launches_value.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1k_launches_activations.csv python bench.py --steps 2 --warmup 3 --profile-activations-leg > gpurun_out/r1k_launches_activations.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_corr9_tma_kernel -c 2 -f -o gpurun_out/r1k_corr9_vgg_conv1 python tools/vgg_bench.py --skip-dense --n-img 376 --reps 1 --layers 1 > /dev/null 2>&1
python tools/vgg_bench.py --reps 2 > gpurun_out/r1k_vgg_full.log 2>&1
python tools/vgg_bench.py --reps 2 --shard 3/8 > gpurun_out/r1k_vgg_shard3of8.log 2>&1
ls -la gpurun_out | tail -8
