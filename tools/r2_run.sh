set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_fc1_launches.csv python tools/dense_bench.py --shapes 25088x4096x1504 --methods auto --reps 1 > gpurun_out/r2c_fc1_ncu.log 2>&1
tail -3 gpurun_out/r2c_fc1_ncu.log
python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,4096x1000x1504 --methods auto --reps 2 2>&1 | tail -5
