set -x
bash scripts/gpu_ab_libs.sh 100m 100000000 ab_fb6 ab_fb8
