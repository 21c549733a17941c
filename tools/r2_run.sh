set -x
timeout 300 python tools/dense_bench.py --shapes 4096x3000x1504,25088x3000x1504 --methods auto --reps 2 2>&1 | grep shape | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lowrank or auto" 2>&1 | tail -3
