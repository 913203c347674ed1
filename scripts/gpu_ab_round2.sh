set -x
python -c "import __graft_entry__ as g; g.build()"
for ub in 0 10 16; do echo "== flat pass, NGPD_UPDATE_BLOCKS=$ub (0 = strategy compiled in, 16 blocks)"; NGPD_UPDATE_BLOCKS=$ub timeout 400 python scripts/gpu_probe_r2.py 100000000 2>&1 | grep chunked | cut -c1-200; done
for ub in 0 10; do
  for st in feature/feature/feature edge/edge/edge; do
    echo "== cubes 10 M, $st, NGPD_UPDATE_BLOCKS=$ub"
    NGPD_UPDATE_BLOCKS=$ub timeout 300 python bench.py --points 10000000 --surface cubes --strategy $st --no-cpu --no-knn --no-extra --no-validate --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}, d['checksum']['pos_hash'])"
  done
done
