set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_e2e_probe.py 10000000 2>&1 | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_knnstats_launches.csv python scripts/gpu_knn_stats.py 4000000 > gpurun_out/ncu_launch_e.log 2>&1; echo "ncu rc=$?"
