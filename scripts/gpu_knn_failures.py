"""Why do queries fall through the 5x5x5 tier?  Pulls the tier-2 hand-over list of the session's last kNN pass and
relates each such query to the grid geometry (run on the GPU box)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import ngpd_b200
from ngpd_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device("cuda:0")
noisy, nrm = bench.make_input(n, dev)
sess = _lib.Session(noisy, 16)
sess.set_state(noisy, nrm)
s, c = sess.mean_edge_length_parts(6)
params = _lib.make_params(16, 8, None, 0.3, 3.0, 0.2, (0, 1, 2), (1.0, 0.2, 1.0), 2.0 * s / c)
sess.step(params)
t0, t1, t2 = sess.knn_stats()
print("tiers", t1, t2, "of", n)
base = _lib.load().ngpd_session_buffer(sess._h, 7)
buf = (ctypes.c_int32 * (3 * n + 3)).from_address  # device memory: copy through torch instead
lists = torch.empty(2 * n + 2, dtype=torch.int32, device=dev)
import ctypes as C
cudart = C.CDLL("libcudart.so")
cudart.cudaMemcpy(C.c_void_p(lists.data_ptr()), C.c_void_p(base), C.c_size_t(4 * (3 * n + 3)), 3)
torch.cuda.synchronize()
l2 = lists[2 * n:2 * n + t2].long()
perm = sess.order().long()                       # tree row -> original index
orig = perm[l2]
g = _lib.Grid(noisy, 16)
gi = g.info()
h = gi.cell_size
lo = torch.tensor(list(gi.bbox)[:3], device=dev, dtype=torch.float64)
dims = torch.tensor(list(gi.dims), device=dev)
print("cell", h, "dims", list(gi.dims), "points per occupied cell", n / gi.occupied_cells)
pos, _, lab = sess.get_state(True)               # positions after the step; the search used the ones before
q = noisy[orig].double()
idx, d2 = g.knn(noisy[orig], 16, _lib.KNN_EXACT_ONLY, with_d2=True)
rk = d2[:, 15].double().sqrt()
cell = ((q - lo) / h).floor().long().clamp_(min=0)
cell = torch.minimum(cell, (dims - 1).long())
frac = (q - lo) / h - cell
reach = torch.full((len(orig),), 1e30, device=dev, dtype=torch.float64)
for a in range(3):
    lo_ok = cell[:, a] - 2 > 0
    hi_ok = cell[:, a] + 2 < dims[a] - 1
    reach = torch.where(lo_ok, torch.minimum(reach, (frac[:, a] + 2) * h), reach)
    reach = torch.where(hi_ok, torch.minimum(reach, (3 - frac[:, a]) * h), reach)
print("k-th distance / h: quantiles", torch.quantile(rk / h, torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99], device=dev, dtype=torch.float64)).tolist())
print("fraction with k-th distance beyond the 5x5x5 reach:", float((rk >= reach).double().mean()))
near = (rk < reach)
print("of the rest (should have been answered):", int(near.sum()))
if near.any():
    qq = q[near][:10]
    print("examples", qq.tolist(), (rk[near][:10] / h).tolist(), cell[near][:10].tolist())
    # near-tie between the 16th and 17th neighbour?
    idx17, d17 = g.knn(noisy[orig][near], 17, _lib.KNN_EXACT_ONLY, with_d2=True)
    gap = (d17[:, 16] - d17[:, 15]).double() / (h * h)
    print("gap 17th-16th in h^2 units: quantiles", torch.quantile(gap, torch.tensor([0.1, 0.5, 0.9], device=dev, dtype=torch.float64)).tolist())
# where are the failing queries?  label histogram and distance to the cube's creases
print("labels of the failing rows", torch.bincount(lab[orig].long(), minlength=3).tolist())
a = q.abs()
on_face = (a.max(dim=1).values - 1).abs() < 0.01
print("near a cube face:", float(on_face.double().mean()), " near a cube edge:", float(((a > 0.99).sum(1) >= 2).double().mean()))
