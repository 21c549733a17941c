set -x
mkdir -p gpurun_out
timeout 900 python tools/grid_bench.py > gpurun_out/r2_grid_bench.log 2>&1; tail -12 gpurun_out/r2_grid_bench.log
