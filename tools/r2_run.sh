set -x
for o in "" "sweep_nt=16"; do
  echo "== opt: $o"
  timeout 300 python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504 --methods auto --reps 2 --opt "$o" 2>&1 | grep shape | cut -c1-150
done
