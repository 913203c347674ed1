set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "knn" > gpurun_out/pytest_q3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q3.log
tail -12 gpurun_out/pytest_q3.log
timeout 900 python scripts/bench_knn_sweep.py 100000000 > gpurun_out/knn_sweep.md 2> gpurun_out/knn_sweep.err; echo "sweep rc=$?"; cat gpurun_out/knn_sweep.md; tail -3 gpurun_out/knn_sweep.err
