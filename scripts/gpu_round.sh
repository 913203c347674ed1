# One full evidence round on ONE B200: GPU tests, smoke, the bench lines (default 100 M incl. the secondary configurations, reference
# arm, 10 M), the ncu launch list and one --set full capture of a steady-state iteration, the one-GPU k-NN / Chamfer sweep.
# Outputs under gpurun_out/ (*_<tag>.*); turn them into profiles/ with scripts/make_profiles.py.
#   gpurun --timeout 2400 -- bash scripts/gpu_round.sh r2a
tag=${1:-r2}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -14 gpurun_out/pytest_$tag.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "orientation_vs_reference or cpsd_loop or ours_clamp or step_at_10M" 2>&1 | grep -E "orientation agrees|CPSD iteration|Ours iteration|10 M-point|passed|failed" > gpurun_out/pytest_prints_$tag.log; cat gpurun_out/pytest_prints_$tag.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_$tag.log
( time timeout 900 python bench.py ) > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_$tag.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_${tag}_10m.json 2> gpurun_out/bench_${tag}_10m.err; echo "bench10 rc=$?"
timeout 300 python scripts/bench_small_configs.py > gpurun_out/small_configs_$tag.md 2> gpurun_out/small_configs_$tag.err; echo "small rc=$?"; cat gpurun_out/small_configs_$tag.md; tail -3 gpurun_out/small_configs_$tag.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --no-extra --no-validate --points 10000000 > gpurun_out/ncu_plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --no-extra --no-validate --points 10000000 > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_' -s 64 -c 18 -o gpurun_out/${tag}_session -f python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --no-extra --no-validate --points 10000000 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 900 python scripts/bench_knn_sweep.py 100000000 > gpurun_out/knn_sweep_$tag.md 2> gpurun_out/knn_sweep_$tag.err; echo "sweep rc=$?"; cat gpurun_out/knn_sweep_$tag.md
python - <<PY
import json
for f in ('gpurun_out/bench_$tag.json','gpurun_out/bench_${tag}_10m.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no line', e); continue
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], 'launches', d['gpu_launches'], 'cold', d['cold'])
    for k,v in d['kernels'].items(): print(' ',k, round(v['ms_per_step'],3),'ms', round(v.get('frac',0),4))
    print(' validated', d['validated']); print(' checksum', d['checksum']); print(' cpu', d['cpu_baseline']); print(' knn', d['knn'])
    for e in d.get('extra_configs') or []: print(' extra', e['name'], round(e['ms_per_step'],3), 'ms', e['class_histogram'], e['kernels_ms_per_step'])
PY
cut -c1-400 gpurun_out/bench_${tag}_ref.json
