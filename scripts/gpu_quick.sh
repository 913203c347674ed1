# quick single-GPU check after a kernel change: the k-NN / session parity tests (without the three slowest cases) and the per-step probe
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "(knn and not clusters) or rerank or session or denoise or phase_driver or run_host or ours or k32 or pruned or update_steps or teacher" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
timeout 200 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_quick.log 2>&1; grep chunked gpurun_out/probe_quick.log | cut -c1-250
NGPD_DELTA_FULL_PASS=1 timeout 200 python scripts/gpu_probe_r2.py 10000000 2>&1 | grep chunked | cut -c1-250 | tail -3
timeout 300 python scripts/bench_small_configs.py 2>&1 | head -5 | cut -c1-300
