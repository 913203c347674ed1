# bench line on N GPUs only:  gpurun --gpus N -- bash scripts/gpu_bench_n.sh N tag [extra bench args]
N=$1; tag=$2; shift; shift
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
if [ "$N" = "1" ]; then
  timeout 900 python bench.py "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
fi
echo "bench rc=$?"; tail -3 gpurun_out/bench_${tag}.err | cut -c1-300
python - <<PY
import json
f='gpurun_out/bench_${tag}.json'
d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'], 'cold', d['cold']['value'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v.get('frac',0),4))
print(' outside', d['roofline']['outside_kernels_ms_per_step'], 'halo', d['halo'], 'tiers', d['knn_tiers_last_step'])
if d.get('kernel_ms_per_step_by_rank'):
    for k,v in d['kernel_ms_per_step_by_rank'].items(): print('  by rank', k, v)
print(' checksum', d['checksum']['pos_hash'], d['checksum']['nrm_hash'], 'validated ok', d['validated'] and d['validated']['ok'])
PY
