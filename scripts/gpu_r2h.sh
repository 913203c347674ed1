set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "rerank or session or denoise or phase_driver or run_host or ours or k32 or teacher" > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2h.log
for m in default ldg rows; do
  NGPD_RERANK=$m timeout 200 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2h_$m.log 2>&1; echo "$m rc=$?"; grep chunked gpurun_out/probe_r2h_$m.log | tail -4 | cut -c1-120
done
