# one GPU call: A/B of the library variants at 100 M and 10 M points + the k = 64 grid variant
set -x
bash scripts/gpu_ab_libs.sh 100m 100000000 ab_old ab_r1g8
bash scripts/gpu_ab_libs.sh 10m 10000000 ab_old
for v in libngpd libngpd_ab_k64; do
  NGPD_LIBRARY=$PWD/normal-guided-pointcloud-denoiser_b200/$v.so timeout 200 python scripts/gpu_probe_k64.py 10000000 64 2>&1 | tail -1
  NGPD_LIBRARY=$PWD/normal-guided-pointcloud-denoiser_b200/$v.so timeout 200 python scripts/gpu_probe_k64.py 10000000 32 2>&1 | tail -1
done
