set -x
mkdir -p gpurun_out
NGPD_NO_TAIL_OVERLAP=1 timeout 300 python scripts/gpu_knn_stats.py 10000000 32 2>&1 | tail -8
timeout 300 python scripts/gpu_knn_stats.py 10000000 32 2>&1 | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:session_ -c 200 --csv --log-file gpurun_out/r1r_k32_launches.csv python scripts/gpu_knn_stats.py 10000000 32 > gpurun_out/ncu_launch_r.log 2>&1; echo "ncu rc=$?"
python scripts/launch_list_md.py gpurun_out/r1r_k32_launches.csv 1 2>&1 | tail -40
