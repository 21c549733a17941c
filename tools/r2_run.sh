set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r2r_bench_vgg.json 2> gpurun_out/r2r_bench_vgg.err; tail -3 gpurun_out/r2r_bench_vgg.err
for w in cifar10_cnn mnist_mlp; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e > gpurun_out/r2r_bench_$w.json 2> gpurun_out/r2r_bench_$w.err; tail -3 gpurun_out/r2r_bench_$w.err; done
python - <<'PY'
import json
for n in ('vgg','cifar10_cnn','mnist_mlp'):
    try:
        l=json.loads(open(f'gpurun_out/r2r_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, {k:l[k] for k in ('value','ms_per_step','gpu_launches')})
        print({k:(round(v['ms'],3), v.get('frac')) for k,v in l['per_layer'].items()})
        print('parity', {k:v.get('agreement') for k,v in l.get('parity',{}).items()} if isinstance(l.get('parity'),dict) else l.get('parity'))
    except Exception as e: print(n, 'ERR', e)
PY
