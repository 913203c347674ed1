"""Which rows reach the exact (last) k-NN tier, and why: their k-th distance in cells, points in their 5x5x5 block, slots needed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import ngpd_b200
from ngpd_b200 import _lib
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
args = argparse.Namespace(surface="creased", strategy="flat/edge/feature", clamp=False)
noisy, analytic, _ = bench.make_shard(args, n, dev, 0, 1)
nrm = bench.single_gpu_normals(noisy, analytic)
g = _lib.Grid(noisy, 16)
gi = g.info()
h = gi.cell_size
lo = torch.tensor(list(gi.bbox)[:3], device=dev, dtype=torch.float64)
print("cell", h, "dims", list(gi.dims), "occupied", gi.occupied_cells, "points/occupied cell", n / gi.occupied_cells, "rebuilds", gi.rebuilds)
sess = _lib.Session(noisy, 16)
sess.reserve(16)
sess.set_state(noisy, nrm)
s, c = sess.mean_edge_length_parts(6)
print("mean 6-NN edge (incl self)", s / c, "in cells", s / c / h)
params = _lib.make_params(dmax=2.0 * s / c)
lib = _lib.load()
cell = ((noisy.double() - lo) / h).floor().long()
for it in range(3):
    pos_before, _, _ = sess.get_state(False)
    sess.step(params)
    torch.cuda.synchronize()
    st = sess.knn_stats()
    fix = torch.as_tensor(bench._DevView(lib.ngpd_session_buffer(sess._h, 7), (3 * n + 3,), "<i4"), device=dev)
    rows_tree = fix[2 * n: 2 * n + st[2]].long().clone()
    perm = sess.order().long()
    rows = perm[rows_tree]
    print(f"it {it}: tiers {st}; rows in the exact tier: {rows.numel()}")
    q = pos_before[rows]
    qc = ((q.double() - lo) / h).floor().long()
    for i in range(min(rows.numel(), 12)):
        d = (noisy - q[i]).double().norm(dim=1)
        dk = torch.topk(d, 33, largest=False).values
        dc = (cell - qc[i]).abs().max(dim=1).values
        in5 = int((dc <= 2).sum()); in3 = int((dc <= 1).sum())
        # slots: rows (dy, dz) of the 5x5x5 block, each split at a brick boundary, 64 points per slot
        sel = dc <= 2
        rel = cell[sel] - qc[i]
        rowid = (rel[:, 2] + 2) * 5 + (rel[:, 1] + 2)
        brick = (cell[sel][:, 0] >> 3)
        key = rowid * 4 + (brick - brick.min())
        cnt = torch.bincount(key)
        slots = int(((cnt + 63) // 64).sum())
        frac = ((q[i].double() - lo) / h) - qc[i]
        print(f"   row {int(rows[i])} pos {[round(float(v),4) for v in q[i]]} d16={float(dk[15])/h:.3f} d32={float(dk[31])/h:.3f} cells; in 3^3: {in3} in 5^3: {in5} slots {slots}; "
              f"pos in cell {[round(float(v),2) for v in frac]}")
