set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_network.py tests/test_gpu_sample_split.py -x -q -m gpu -k "conv or corr or cnn or vgg or gram_i8_non_finite" 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r2q_bench_vgg.json 2> gpurun_out/r2q_bench_vgg.err; tail -3 gpurun_out/r2q_bench_vgg.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2q_bench_vgg.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')})
print({k:(v['ms'], v.get('frac')) for k,v in l['per_layer'].items()})
PY
