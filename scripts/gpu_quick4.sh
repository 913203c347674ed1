set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "orientation or division or mesh or preprocess" > gpurun_out/pytest_q4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q4.log
tail -25 gpurun_out/pytest_q4.log
python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import ngpd_b200 as ng
from ngpd_b200 import workloads
for n in (1_000_000, 10_000_000):
    clean, normal = workloads.creased_surface(n, 1234, 'cuda')
    noisy = workloads.add_noise(clean, 0.3 * workloads.expected_spacing(n))
    p = ng.Processor(ng.Pointcloud(noisy))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p.graph.edge_index = p.graphBuilder.getKNNEdgeIndex(12)
    p.graphBuilder.setAndFlipNormals(flip=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    p.graphBuilder.flipNormals()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    agree = float(((p.graph.n * normal).sum(1) > 0).float().mean())
    print(f"n={n}: kNN(12)+PCA {1e3*(t1-t0):.1f} ms, orientation {1e3*(t2-t1):.1f} ms {p.graphBuilder.orientation_info}, agreement with the analytic normals {max(agree,1-agree):.4%}")
PY
