set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "check rc=$?"
tail -4 gpurun_out/multi_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --points 10000000 > gpurun_out/bench_multi_${N}_10m.json 2> gpurun_out/bench_multi_${N}_10m.err; echo "bench 10M rc=$?"
tail -3 gpurun_out/bench_multi_${N}_10m.err; cut -c1-900 gpurun_out/bench_multi_${N}_10m.json
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --points 100000000 ) > gpurun_out/bench_multi_${N}_100m.json 2> gpurun_out/bench_multi_${N}_100m.err; echo "bench 100M rc=$?"
tail -6 gpurun_out/bench_multi_${N}_100m.err; cut -c1-900 gpurun_out/bench_multi_${N}_100m.json
