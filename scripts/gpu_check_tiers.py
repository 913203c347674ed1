"""Full-table check of the session's tiered k-NN (re-ranking + streaming tiers with stored candidates) against the exact shell search,
iteration by iteration on the bench cloud: two sessions in lockstep, one per mode, neighbour tables compared row by row."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import ngpd_b200
from ngpd_b200 import _lib
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
its = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
args = argparse.Namespace(surface="creased", strategy="flat/edge/feature", clamp=False)
noisy, analytic, _ = bench.make_shard(args, n, dev, 0, 1)
nrm = bench.single_gpu_normals(noisy, analytic)
lib = _lib.load()
sess = []
for mode in (0, 1):
    s = _lib.Session(noisy, 16)
    s.set_knn_mode(mode)
    s.set_state(noisy, nrm)
    sess.append(s)
a, c = sess[0].mean_edge_length_parts(6)
params = _lib.make_params(dmax=2.0 * a / c)
view = lambda s: torch.as_tensor(bench._DevView(lib.ngpd_session_buffer(s._h, 6), (n, 16), "<i4"), device=dev)
for it in range(its):
    for s in sess:
        s.step(params)
    torch.cuda.synchronize()
    bad = (view(sess[0]) != view(sess[1])).any(dim=1)
    pa, na, la = sess[0].get_state(True)
    pb, nb, lb = sess[1].get_state(True)
    print(f"iteration {it}: tiers {sess[0].knn_stats()}; rows that differ from the exact search: {int(bad.sum())}; "
          f"positions equal {bool(torch.equal(pa, pb))}, labels differ {int((la != lb).sum())}", flush=True)
    if bad.any():
        i = int(bad.nonzero()[0])
        print("   first bad tree row", i, view(sess[0])[i].tolist(), view(sess[1])[i].tolist())
        break
