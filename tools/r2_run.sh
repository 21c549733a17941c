set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sweep_lowrank or slgemm" 2>&1 | tail -15
timeout 300 python tools/dense_bench.py --shapes 25088x4096x1504,25088x512x1504,4096x4096x1504,4096x1000x1504 --methods auto --reps 2 2>&1 | grep shape | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_tc_fc1_launches.csv \
    python tools/dense_bench.py --shapes 25088x512x1504 --methods auto --reps 0 > /dev/null 2>&1
