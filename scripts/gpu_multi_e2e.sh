set -x
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -14 gpurun_out/topo.txt | cut -c1-200
lscpu | grep -i "numa\|socket\|^CPU(s)" 
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --points 100000000 ) > gpurun_out/bench_numa_${N}_100m.json 2> gpurun_out/bench_numa_${N}_100m.err; echo "bench 100M rc=$?"
grep "\[bench\]\|real" gpurun_out/bench_numa_${N}_100m.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_numa_${N}_100m.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
PY
