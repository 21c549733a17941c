set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_fullsize.log; cat gpurun_out/r2_fullsize.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_vgg.json 2> gpurun_out/r2a_bench_vgg.err; tail -c 3000 gpurun_out/r2a_bench_vgg.json; tail -5 gpurun_out/r2a_bench_vgg.err
