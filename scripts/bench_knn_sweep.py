"""BASELINE.json configs[4]: kNN-only and Chamfer-only sweep on the synthetic creased surface, one B200.
    python scripts/bench_knn_sweep.py [max_points]   -> markdown table on stdout (profiles/r1_knn_sweep.md)
kNN = ngpd_knn (public ABI: every query searched, int32 index table materialised); Chamfer = two ngpd_nn_sqdist passes
between the noisy cloud and its clean surface, grids built beforehand (build time listed separately)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ngpd_b200
from ngpd_b200 import _lib, workloads

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda:0")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print(f"| points | k | grid build ms | kNN ms | G queries/s | algorithmic GB/s | frac of {peak:.0f} GB/s | Chamfer (both directions) ms | G points/s |")
print("|---|---|---|---|---|---|---|---|---|")
for n in (1_000_000, 10_000_000, 100_000_000):
    if n > nmax:
        break
    clean, _ = workloads.creased_surface(n, 1234, dev)
    noisy = workloads.add_noise(clean, 0.3 * workloads.expected_spacing(n))
    for k in (8, 16, 32, 64):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        grid = _lib.Grid(noisy, k)
        torch.cuda.synchronize(); build = (time.perf_counter() - t0) * 1e3
        out = {}
        ms = timed(lambda: out.__setitem__("t", grid.knn(noisy, k, _lib.KNN_QUERY_IS_TREE)))
        out.clear()
        cham = ""
        if k == 16:
            gc = _lib.Grid(clean, 16)
            cms = timed(lambda: (grid.nn_sqdist(clean, False, 0), gc.nn_sqdist(noisy, False, 0)))
            cham = f"{cms:.2f} | {2 * n / (cms * 1e-3) / 1e9:.2f}"
            del gc
        else:
            cham = "- | -"
        gbs = (24 + 4 * k) * n / (ms * 1e-3) / 1e9
        print(f"| {n} | {k} | {build:.1f} | {ms:.2f} | {n / (ms * 1e-3) / 1e9:.3f} | {gbs:.0f} | {gbs / peak:.3f} | {cham} |", flush=True)
        del grid
    del clean, noisy
    torch.cuda.empty_cache()
