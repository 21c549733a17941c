set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2s_bench_vgg_n8.json 2> gpurun_out/r2s_bench_vgg_n8.err; tail -3 gpurun_out/r2s_bench_vgg_n8.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2s_bench_vgg_n8.json').read().strip().splitlines()[-1])
print({k:l.get(k) for k in ('value','ms_per_step','gpu_launches','n_gpus')}); print(l.get('e2e',{}).get('ms_per_step'))
print({k:(round(v['ms'],3), v.get('frac')) for k,v in l['per_layer'].items()})
PY
