set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2t_bench_vgg.json 2> gpurun_out/r2t_bench_vgg.err; tail -3 gpurun_out/r2t_bench_vgg.err
python - <<'PY'
import json
for n in ('vgg',):
    l=json.loads(open(f'gpurun_out/r2t_bench_{n}.json').read().strip().splitlines()[-1])
    print(n, {k:l[k] for k in ('value','ms_per_step','gpu_launches')})
    print({k:(round(v['ms'],3), v.get('frac')) for k,v in l['per_layer'].items()})
PY
