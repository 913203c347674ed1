set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_r2e.log 2>&1; echo "check rc=$?"; grep -v "^W1018\|^\*\*\*\|^$" gpurun_out/multi_check_r2e.log | head -30
