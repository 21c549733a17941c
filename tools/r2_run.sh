mkdir -p gpurun_out
python tools/dense_methods.py --reps 2 --out gpurun_out/dense_methods_r2.md 2>&1 | tail -25
python tools/grid_bench.py 2>&1 | tail -2
