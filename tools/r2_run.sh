set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_network.py tests/test_gpu_fullsize.py -x -q -m gpu -k "conv or corr or cnn" 2>&1 | tail -4
