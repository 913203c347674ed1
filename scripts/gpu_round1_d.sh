set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_e2e_probe.py 10000000 2>&1 | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'session_knn_(rerank|fast|wide)' -s 6 -c 5 -o gpurun_out/r1f_rerank -f python scripts/gpu_knn_stats.py 4000000 > gpurun_out/ncu_full_f.log 2>&1; echo "ncu full rc=$?"
