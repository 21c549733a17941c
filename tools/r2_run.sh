set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slgemm" 2>&1 | tail -30 > gpurun_out/r2b_slgemm.log; cat gpurun_out/r2b_slgemm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lowrank" 2>&1 | tail -30 > gpurun_out/r2b_lowrank.log; cat gpurun_out/r2b_lowrank.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "fc1 or fc2" 2>&1 | tail -30 > gpurun_out/r2b_fullsize.log; cat gpurun_out/r2b_fullsize.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r2b_bench_vgg.json 2> gpurun_out/r2b_bench_vgg.err; tail -c 1200 gpurun_out/r2b_bench_vgg.json; tail -5 gpurun_out/r2b_bench_vgg.err
