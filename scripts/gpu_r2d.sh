# 2 GPUs: slab == single-GPU check (both plans, both transports, clamp mode, narrow-halo detection), then the bench on slabs
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_r2d.log 2>&1; echo "check rc=$?"; grep -v "^W\|^\*\*\*" gpurun_out/multi_check_r2d.log | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --points 10000000 --steps 8 --warmup 3 > gpurun_out/bench_r2d_2gpu_10m.json 2> gpurun_out/bench_r2d_2gpu_10m.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2d_2gpu_10m.err; cut -c1-2500 gpurun_out/bench_r2d_2gpu_10m.json
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu --no-knn > gpurun_out/bench_r2d_1gpu_10m.json 2> gpurun_out/bench_r2d_1gpu_10m.err; echo "bench1 rc=$?"; tail -3 gpurun_out/bench_r2d_1gpu_10m.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2d_1gpu_10m.json','gpurun_out/bench_r2d_2gpu_10m.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no line', e); continue
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'], 'cold', d['cold'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v.get('frac',0),4))
    print(' outside', d['roofline']['outside_kernels_ms_per_step'], 'checksum', d['checksum'], 'validated', d['validated'], 'halo', d['halo'])
PY
