set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rerank or session or fast_path or denoise or until or cube" -s > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log | cut -c1-300
timeout 300 python scripts/gpu_knn_stats.py 4000000 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
PY
