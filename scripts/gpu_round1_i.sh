set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_i.log
tail -5 gpurun_out/pytest_i.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 46 -c 16 -o gpurun_out/r1i_session -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_i.log 2>&1; echo "ncu full rc=$?"
