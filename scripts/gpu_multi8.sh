set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "check rc=$?"
tail -3 gpurun_out/multi_check_$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 ) > gpurun_out/bench_multi_${N}_100m.json 2> gpurun_out/bench_multi_${N}_100m.err; echo "bench 100M rc=$?"
tail -6 gpurun_out/bench_multi_${N}_100m.err; cut -c1-300 gpurun_out/bench_multi_${N}_100m.json
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 8 --warmup 3 ) > gpurun_out/bench_multi_4_100m.json 2> gpurun_out/bench_multi_4_100m.err; echo "bench 4 100M rc=$?"
tail -6 gpurun_out/bench_multi_4_100m.err; cut -c1-300 gpurun_out/bench_multi_4_100m.json
