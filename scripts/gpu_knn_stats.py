"""kNN tier statistics and per-kernel timing of the session on the bench workload (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import ngpd_b200
from ngpd_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
K_F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
noisy, nrm = bench.make_input(n, dev)
sess = _lib.Session(noisy, K_F)
sess.set_state(noisy, nrm)
s, c = sess.mean_edge_length_parts(6)
print("6-NN pass tiers:", sess.knn_stats(), "of", n)
params = _lib.make_params(K_F, 8, None, 0.3, 3.0, 0.2, (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE), (1.0, 0.2, 1.0), 2.0 * s / c)
gi = _lib.GridInfo(); 
for it in range(4):
    sess.set_profiling(True); sess.get_profile()
    sess.step(params)
    torch.cuda.synchronize()
    prof = sess.get_profile()
    print(f"iteration {it}: tiers {sess.knn_stats()}  " + "  ".join(f"{k} {v[0]:.3f} ms" for k, v in prof.items()))
