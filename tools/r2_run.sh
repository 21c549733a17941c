mkdir -p gpurun_out
for o in "" "sweep_range=1024" "sweep_range=1024,sweep_groups=1" "sweep_range=768" "sweep_range=2048"; do
  echo "== opt: $o"
  python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,4096x1000x1504,25088x512x1504 --methods auto --reps 2 --opt "$o" 2>&1 | grep shape | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lowrank or golden or kat or dense" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "fc1 or fc2" 2>&1 | tail -3
