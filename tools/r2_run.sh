set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_vgg16_n1.json 2> gpurun_out/r2_bench_vgg16_n1.err
python bench.py --workload cifar10_cnn --steps 10 --warmup 3 > gpurun_out/r2_bench_cifar10_cnn_n1.json 2> /dev/null
python bench.py --workload mnist_mlp --steps 10 --warmup 3 > gpurun_out/r2_bench_mnist_mlp_n1.json 2> /dev/null
python - <<'PY'
import json
for n in ('vgg16_n1','cifar10_cnn_n1','mnist_mlp_n1'):
    l=json.loads(open(f'gpurun_out/r2_bench_{n}.json').read().strip().splitlines()[-1])
    print(n, round(l['ms_per_step'],3), l['e2e']['ms_per_step'], {k:v.get('walk') for k,v in l['per_layer'].items() if 'walk' in v})
PY
