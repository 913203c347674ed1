set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 300 python scripts/gpu_probe_r2.py 100000000 > gpurun_out/probe_100m.log 2>&1; grep chunked gpurun_out/probe_100m.log | cut -c1-220
NGPD_NO_TAIL_OVERLAP=1 timeout 300 python scripts/gpu_probe_r2.py 100000000 > gpurun_out/probe_100m_notail.log 2>&1; grep chunked gpurun_out/probe_100m_notail.log | cut -c1-220
for ub in 10 8 6; do NGPD_UPDATE_BLOCKS=$ub timeout 200 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_update_$ub.log 2>&1; echo "update blocks $ub"; grep chunked gpurun_out/probe_update_$ub.log | tail -3 | cut -c1-130; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'session_knn|session_nvt_smooth' -c 60 --csv --log-file gpurun_out/knn_launches_100m.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --no-extra --no-validate > gpurun_out/ncu_knn_100m.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/knn_launches_100m.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki,vi,ui=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit')
for r in rows[start+1:]:
    if len(r)>vi: print(r[ki].split('(')[0].replace('void ','').replace('ngpd::','')[:60], r[vi], r[ui])
PY
