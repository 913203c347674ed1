# Slab runs on N GPUs of one box: bit-identity check, the bench line at 100 M points, optionally the sweep.
#   gpurun --gpus N --timeout 1500 -- bash scripts/gpu_scale.sh N tag [sweep_max_points]
N=$1; tag=$2; sweep=${3:-0}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_${tag}.log 2>&1; echo "check rc=$?"; grep -v "^W1018\|^\*\*\*\|^$" gpurun_out/multi_check_${tag}.log | grep "^\[\|MULTI" | cut -c1-400
timeout 900 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${tag}.err | cut -c1-300
python - <<PY
import json
f='gpurun_out/bench_${tag}.json'
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'], 'cold', d['cold']['value'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v.get('frac',0),4))
    print(' outside', d['roofline']['outside_kernels_ms_per_step'], 'halo', d['halo']); print(' validated', d['validated']); print(' checksum', d['checksum'])
except Exception as e:
    print('no bench line', e)
PY
if [ "$sweep" != "0" ]; then
  timeout 1200 $TR --master-port 29513 scripts/bench_knn_sweep.py $sweep 100000000 > gpurun_out/knn_sweep_${tag}.md 2> gpurun_out/knn_sweep_${tag}.err; echo "sweep rc=$?"; cat gpurun_out/knn_sweep_${tag}.md; grep -v "^W1018\|^\*\*\*\|^$" gpurun_out/knn_sweep_${tag}.err | tail -8 | cut -c1-300
fi
