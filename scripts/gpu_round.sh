# One full evidence round on a B200 box: GPU tests, smoke, the bench lines (default 100 M, reference arm, 10 M, 10 M k=32), small configs,
# the ncu launch list and one --set full capture of a steady-state iteration.  Outputs under gpurun_out/ (*_r.*, rN_*); turn them into
# profiles/ with scripts/make_profiles.py.   gpurun --timeout 1500 -- bash scripts/gpu_round.sh
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -4 gpurun_out/pytest_r.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_r.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_r.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_r.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_r_ref.json 2> gpurun_out/bench_r_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_r_10m.json 2> gpurun_out/bench_r_10m.err; echo "bench10 rc=$?"
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu --k-feature 32 > gpurun_out/bench_r_10m_k32.json 2> gpurun_out/bench_r_10m_k32.err; echo "bench10 k32 rc=$?"
timeout 300 python scripts/bench_small_configs.py > gpurun_out/small_configs.md 2> gpurun_out/small_configs.err; echo "small rc=$?"; cat gpurun_out/small_configs.md; tail -3 gpurun_out/small_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/rN_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --points 10000000 > gpurun_out/ncu_launch_r.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 50 -c 18 -o gpurun_out/rN_session -f python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --points 10000000 > gpurun_out/ncu_full_r.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_r.json','gpurun_out/bench_r_10m.json','gpurun_out/bench_r_10m_k32.json'):
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v['frac'],4))
    print(d['cpu_baseline']); print(d['knn'])
PY
cut -c1-300 gpurun_out/bench_r_ref.json
