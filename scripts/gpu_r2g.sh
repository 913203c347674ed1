set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "rerank or session or denoise or phase_driver or run_host or ours or k32" > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2g.log
timeout 200 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2g_tma.log 2>&1; echo "rc=$?"; grep chunked gpurun_out/probe_r2g_tma.log | tail -5
NGPD_RERANK_LDG=1 timeout 200 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2g_ldg.log 2>&1; grep chunked gpurun_out/probe_r2g_ldg.log | tail -5
