set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lowrank or slgemm" 2>&1 | tail -30 > gpurun_out/r2j_tests.log; cat gpurun_out/r2j_tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "fc1 or fc2" 2>&1 | tail -30 > gpurun_out/r2j_fullsize.log; cat gpurun_out/r2j_fullsize.log
python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,4096x1000x1504,16384x16384x5000 --methods auto --reps 2 --opt sweep_nt=32 2>&1 | grep shape | cut -c1-250
python tools/dense_bench.py --shapes 25088x512x1504,4096x512x1504 --methods auto --reps 2 2>&1 | grep shape | cut -c1-250
