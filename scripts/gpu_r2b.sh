set -x
mkdir -p gpurun_out
timeout 600 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2b.log 2>&1; echo "probe rc=$?"; cat gpurun_out/probe_r2b.log | tail -30
