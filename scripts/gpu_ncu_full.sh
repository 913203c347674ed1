# one plain run, then the --set full capture of the session kernels of one warm step (2M points)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --points 4000000 --no-cpu"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:'session_(knn_fast|knn_fix|nvt_smooth|nvt_classify|update)_kernel' -s 25 -c 7 \
   -o gpurun_out/r1b_session -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
timeout 600 python -m pytest tests -m gpu -x -q -k "free_running or fast_path" -s 2>&1 | tail -15
