mkdir -p gpurun_out
for o in "" "sweep_wq=2" "sweep_wq=2,sweep_groups=2" "sweep_wq=2,sweep_groups=2,sweep_nt=16" "sweep_wq=2,sweep_groups=1"; do
  echo "== opt: $o"
  python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,25088x512x1504 --methods auto --reps 2 --opt "$o" 2>&1 | grep shape | cut -c1-200
done
