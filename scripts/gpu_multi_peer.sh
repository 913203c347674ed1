# halo transport A/B on N GPUs: slab == single-GPU check with the peer-memory push, then bench.py at 10 M points with both transports
set -x
mkdir -p gpurun_out
N=${1:-2}
NGPD_HALO_TRANSPORT=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_peer_$N.log 2>&1; echo "check rc=$?"
tail -5 gpurun_out/multi_check_peer_$N.log
for t in nccl peer; do
  NGPD_HALO_TRANSPORT=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --points ${2:-10000000} --no-knn > gpurun_out/bench_${t}_${N}.json 2> gpurun_out/bench_${t}_${N}.err; echo "bench $t rc=$?"
  tail -2 gpurun_out/bench_${t}_${N}.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${t}_${N}.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$t', 'ms/step', d['ms_per_step'], 'kernel sum', sum(v['ms_per_step'] for v in d['kernels'].values()), 'e2e', d['e2e']['value'])
PY
done
