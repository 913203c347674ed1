set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "cpsd or ball" > gpurun_out/pytest_k_cpsd.log 2>&1; echo "pytest cpsd rc=$?" >> gpurun_out/pytest_k_cpsd.log
tail -25 gpurun_out/pytest_k_cpsd.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k.log
tail -5 gpurun_out/pytest_k.log
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_k_10m.json 2> gpurun_out/bench_k_10m.err; echo "bench10 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_k_10m.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
PY
