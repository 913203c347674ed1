# last check of a build on one B200: the GPU tests, smoke(), the default bench line and the 10 M-point line
#   gpurun --timeout 1500 -- bash scripts/gpu_final.sh <tag>
tag=${1:-final}
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_${tag}_10m.json 2> gpurun_out/bench_${tag}_10m.err; echo "bench10 rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_$tag.json','gpurun_out/bench_${tag}_10m.json'):
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f,'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'cold',d['cold']['value'],'validated',d['validated']['ok'],d['checksum']['pos_hash'],d['checksum']['nrm_hash'])
    print('  ', {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
    for e in d.get('extra_configs') or []: print('   extra', e['name'][:60], round(e['ms_per_step'],3))
PY
