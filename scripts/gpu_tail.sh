echo "== overlap on"; timeout 300 python scripts/gpu_tail_probe.py 10000000 2>&1 | tail -8
echo "== overlap off"; NGPD_NO_TAIL_OVERLAP=1 timeout 300 python scripts/gpu_tail_probe.py 10000000 2>&1 | tail -8
