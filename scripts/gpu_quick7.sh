set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "rerank or session_large or run_host" > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s.log
tail -5 gpurun_out/pytest_s.log
timeout 300 python scripts/gpu_knn_stats.py 10000000 32 2>&1 | tail -5
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu --k-feature 32 > gpurun_out/bench_s_k32.json 2> gpurun_out/bench_s_k32.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_s_k32.err
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_s_10m.json 2> gpurun_out/bench_s_10m.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_s_10m.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_s_k32.json','gpurun_out/bench_s_10m.json'):
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
    print(d['knn'])
PY
