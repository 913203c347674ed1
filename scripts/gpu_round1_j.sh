set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_j.log
tail -4 gpurun_out/pytest_j.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_j.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_j.log
( time timeout 900 python bench.py ) > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_j.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_j_ref.json 2> gpurun_out/bench_j_ref.err; echo "bench ref rc=$?"; tail -4 gpurun_out/bench_j_ref.err
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_j_10m.json 2> gpurun_out/bench_j_10m.err; echo "bench10 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --points 10000000 > gpurun_out/ncu_launch_j.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 50 -c 18 -o gpurun_out/r1j_session -f python bench.py --steps 2 --warmup 3 --no-cpu --points 10000000 > gpurun_out/ncu_full_j.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_j.json','gpurun_out/bench_j_10m.json'):
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'], d['roofline'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
    print(d['cpu_baseline'])
PY
cut -c1-600 gpurun_out/bench_j_ref.json
