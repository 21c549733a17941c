set -x
K='regex:conv_gram|conv_fin|conv_sweep|gemm_nt|sweep_|dense_stream|row_norms|transpose_|reduce_splits|msq_|gram_i8|i8_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1g_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -c 1 -f -o gpurun_out/r1g_i8_dense19 python tools/dense_bench.py --shapes 2048x128x5008 --methods gram_i8 --reps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -c 1 -f -o gpurun_out/r1g_i8_4096 python tools/dense_bench.py --shapes 4096x4096x25000 --methods gram_i8 --reps 0 > /dev/null 2>&1
ls -la gpurun_out | tail -8
