set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -5 gpurun_out/pytest_g.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_g_ref.json 2> gpurun_out/bench_g_ref.err; echo "bench ref rc=$?"
( time timeout 900 python bench.py --steps 5 --warmup 3 --points 100000000 --no-cpu ) > gpurun_out/bench_g_100m.json 2> gpurun_out/bench_g_100m.err; echo "bench 100M rc=$?"
tail -5 gpurun_out/bench_g_100m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_g.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 46 -c 12 -o gpurun_out/r1g_session -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_g.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_g.json','gpurun_out/bench_g_100m.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
    print(d['cpu_baseline'])
PY
cat gpurun_out/bench_g_ref.json
