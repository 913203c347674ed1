"""k-NN rate of the public entry (ngpd_knn, every query searched) for one k on the bench cloud: python scripts/gpu_probe_k64.py [points] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import ngpd_b200
from ngpd_b200 import _lib
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
args = argparse.Namespace(surface="creased", strategy="flat/edge/feature", clamp=False)
noisy, _, _ = bench.make_shard(args, n, dev, 0, 1)
grid = _lib.Grid(noisy, k)
out = grid.knn(noisy, k, _lib.KNN_QUERY_IS_TREE)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    out = grid.knn(noisy, k, _lib.KNN_QUERY_IS_TREE)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"k={k} n={n}: {ms:.2f} ms, {n / ms / 1e6:.3f} G queries/s, row hash {int(out.long().sum())}", flush=True)
