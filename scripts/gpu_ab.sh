# A/B timing of alternative builds of libngpd.so (NGPD_LIBRARY) on the bench workload: per-kernel-group ms per iteration
mkdir -p gpurun_out
D=$PWD/normal-guided-pointcloud-denoiser_b200
for v in "" $@; do
  lib=$D/libngpd${v:+_$v}.so
  for k in 16 32; do
    echo "== ${v:-main} k=$k"
    NGPD_LIBRARY=$lib timeout 300 python scripts/gpu_knn_stats.py 10000000 $k 2>&1 | tail -2
  done
done
