set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --points 4000000 --no-cpu"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:'session_(knn_fast|knn_wide|knn_fix|nvt_smooth)_kernel' -s 16 -c 4 \
   -o gpurun_out/r1c_knn -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q -k "until_min or generic_strat or cube or preprocess or session_large or eigh3_device" -s 2>&1 | tail -40
