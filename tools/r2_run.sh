set -x
mkdir -p gpurun_out
python bench.py --workload cifar10_cnn --steps 10 --warmup 3 > gpurun_out/r2_bench_cifar10_cnn_n1.json 2> /dev/null
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_cifar10_cnn_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['ms_per_step'])
print({k:(round(v['ms'],3), v.get('frac')) for k,v in l['per_layer'].items()})
print(min(v['agreement'] for v in l['parity'].values()))
PY
