"""Timeline of the overlapped search tail: step time with and without the side stream (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import ngpd_b200
from ngpd_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
noisy, nrm = bench.make_input(n, dev)
sess = _lib.Session(noisy, 16)
sess.set_state(noisy, nrm)
s, c = sess.mean_edge_length_parts(6)
params = _lib.make_params(16, 8, None, 0.3, 3.0, 0.2, (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE), (1.0, 0.2, 1.0), 2.0 * s / c)
for it in range(8):
    sess.set_profiling(True); sess.get_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.step(params); e1.record(); torch.cuda.synchronize()
    prof = sess.get_profile()
    print(f"iteration {it}: step {e0.elapsed_time(e1):.3f} ms tiers {sess.knn_stats()}  " + "  ".join(f"{k} {v[0]:.3f}" for k, v in prof.items()))
