# round 2, first GPU contact: the whole GPU test suite, smoke, a short 10 M-point bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -25 gpurun_out/pytest_r2a.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_r2a.log
timeout 600 python bench.py --points 10000000 --steps 8 --warmup 3 > gpurun_out/bench_r2a_10m.json 2> gpurun_out/bench_r2a_10m.err; echo "bench10 rc=$?"; tail -5 gpurun_out/bench_r2a_10m.err
cut -c1-3000 gpurun_out/bench_r2a_10m.json
