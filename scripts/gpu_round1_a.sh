set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python scripts/gpu_knn_failures.py 4000000 > gpurun_out/knn_failures.log 2>&1; echo "failures rc=$?"
cat gpurun_out/knn_failures.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 30 -c 12 -o gpurun_out/r1d_session -f python bench.py --steps 2 --warmup 3 --points 4000000 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
print(d['cpu_baseline'])
PY
