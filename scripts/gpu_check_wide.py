"""Rows of the tiered public k-NN against the exact shell search, on the queries that reach the 5x5x5 tier (dense regions)."""
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
for k in (13, 16, 32):
    g = _lib.Grid(noisy, k)
    # queries: the dense band around the torus / cube-face intersection plus a random sample
    band = ((noisy[:, 0] - 1.0).abs() < 0.004).nonzero().flatten()[:400000]
    rnd = torch.randperm(n, device=dev)[:200000]
    for name, sel in (("band", band), ("random", rnd)):
        q = noisy[sel].contiguous()
        a = g.knn(q, k, 0)
        b = g.knn(q, k, _lib.KNN_EXACT_ONLY)
        bad = (a != b).any(dim=1)
        print(f"k={k} {name}: {q.size(0)} queries, rows that differ from the exact search: {int(bad.sum())}", flush=True)
        if bad.any():
            i = int(bad.nonzero()[0])
            print("   first:", sel[i].item(), a[i].tolist(), b[i].tolist())
    del g
