set -x
K='regex:conv_|gemm_nt|sweep_|dense_stream|row_norms|transpose_|reduce_splits|msq_|im2col'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1c_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gram9_tma -s 3 -c 3 -f -o gpurun_out/r1c_conv_tma python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_stream_reg -s 1 -c 1 -f -o gpurun_out/r1c_stream python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gram9_nhwc -s 1 -c 3 -f -o gpurun_out/r1c_nhwc python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
