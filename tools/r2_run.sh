set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slgemm or lowrank or gram_i8 or gram_kernels or gram_stage" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python tools/dense_bench.py --shapes 25088x4096x1504,4096x4096x1504,4096x1000x1504,25088x512x1504,16384x16384x5000,4096x4096x25000,2048x128x5008 --methods auto --reps 2 2>&1 | grep shape | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_fc1_launches.csv python tools/dense_bench.py --shapes 25088x4096x1504 --methods auto --reps 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r2p_fc1_launches.csv') if l.startswith('"')))
hdr=rows[0]; ki,vi,gi=hdr.index("Kernel Name"),hdr.index("Metric Value"),hdr.index("Grid Size")
for r in rows[300:310]: print(r[ki].split("(")[0][:40], r[vi], r[gi])
PY
