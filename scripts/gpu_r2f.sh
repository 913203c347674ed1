set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_r2f.log 2>&1; echo "check rc=$?"; grep -v "^W1018\|^\*\*\*\|^$" gpurun_out/multi_check_r2f.log | head -12 | cut -c1-600
CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2f.log 2>&1; grep chunked gpurun_out/probe_r2f.log
CUDA_VISIBLE_DEVICES=0 NGPD_NO_TAIL_OVERLAP=1 timeout 300 python scripts/gpu_probe_r2.py 10000000 > gpurun_out/probe_r2f_notail.log 2>&1; grep chunked gpurun_out/probe_r2f_notail.log
