set -x
timeout 300 python scripts/gpu_probe_fix_rows.py 10000000 > gpurun_out/probe_fix_rows.log 2>&1; echo rc=$?; tail -50 gpurun_out/probe_fix_rows.log
