set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rerank or session_large or run_host or phase" > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -3 gpurun_out/pytest_q.log
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_q.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
print(d['knn'])
PY
timeout 900 python scripts/bench_knn_sweep.py 100000000 > gpurun_out/knn_sweep.md 2> gpurun_out/knn_sweep.err; echo "sweep rc=$?"; cat gpurun_out/knn_sweep.md; tail -3 gpurun_out/knn_sweep.err
