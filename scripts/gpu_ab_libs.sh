# A/B of library variants built next to the default one (build.py: NGPD_LIB_OUT / NGPD_EXTRA_NVCC_FLAGS), same box, same input:
#   bash scripts/gpu_ab_libs.sh <tag> <points> <lib suffix> [<lib suffix> ...]     ("" = the default libngpd.so)
set -x
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
for v in default "$@"; do
  lib=$PWD/normal-guided-pointcloud-denoiser_b200/libngpd_$v.so
  [ "$v" = default ] && lib=$PWD/normal-guided-pointcloud-denoiser_b200/libngpd.so
  NGPD_LIBRARY=$lib timeout 400 python scripts/gpu_probe_r2.py $n > gpurun_out/ab_${tag}_$v.log 2>&1
  echo "== $v"; grep chunked gpurun_out/ab_${tag}_$v.log | cut -c1-260
done
