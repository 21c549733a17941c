set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "sweep or lowrank or tensor_core or fc1 or fc2 or auto or dense" 2>&1 | tail -6
timeout 300 python tools/dense_bench.py --shapes 25088x4096x1504,25088x512x1504,4096x512x1504,4096x125x1504,4096x4096x1504 --methods auto --reps 2 2>&1 | grep shape | cut -c1-200
