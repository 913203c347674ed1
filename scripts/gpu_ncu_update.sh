set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 170 ncu --set full --clock-control none --import-source on -k regex:'session_update_kernel|session_knn_rerank_kernel' -s 8 -c 4 -o gpurun_out/r2g_update -f python bench.py --steps 2 --warmup 3 --no-cpu --no-knn --no-extra --no-validate --points 10000000 > gpurun_out/ncu_update_r2g.log 2>&1; echo "ncu rc=$?"
