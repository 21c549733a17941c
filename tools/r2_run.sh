set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lowrank or dense or golden or kat or shard" 2>&1 | tail -30 > gpurun_out/r2i_tests.log; cat gpurun_out/r2i_tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "fc1 or fc2 or dense19 or config4_4096" 2>&1 | tail -30 > gpurun_out/r2i_fullsize.log; cat gpurun_out/r2i_fullsize.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_fc1_launches.csv python tools/dense_bench.py --shapes 25088x4096x1504 --methods auto --reps 1 > gpurun_out/r2i_fc1_ncu.log 2>&1
python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,4096x1000x1504,2048x128x5008,4096x4096x25000 --methods auto --reps 2 2>&1 | tail -5
