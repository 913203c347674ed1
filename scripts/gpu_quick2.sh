set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in a b; do timeout 200 python scripts/gpu_probe_r2.py 10000000 2>&1 | grep chunked | cut -c95-260 | head -4; done
NGPD_TAIL_OVERLAP=1 timeout 200 python scripts/gpu_probe_r2.py 10000000 2>&1 | grep chunked | cut -c95-260 | head -4
NGPD_RERANK=ldg timeout 200 python scripts/gpu_probe_r2.py 10000000 2>&1 | grep chunked | cut -c95-260 | head -4
NGPD_NO_FAST_LABELS=1 timeout 200 python scripts/gpu_probe_r2.py 10000000 2>&1 | grep chunked | cut -c95-260 | head -4
