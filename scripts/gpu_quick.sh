set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rerank or session or denoise or until or labels or generic or run_host or phase or one_shot" > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -5 gpurun_out/pytest_q.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
PY
