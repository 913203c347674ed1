"""BASELINE.json configs[4]: kNN-only and Chamfer-only sweep on the synthetic creased surface, 1 M - 1 B points, k in {8,16,32,64},
on one B200 or (under torchrun) on the GPUs of one box.

    python scripts/bench_knn_sweep.py [max_points] [min_points]                                     # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/bench_knn_sweep.py 1000000000   # slabs

One GPU: kNN = ngpd_knn (public ABI: every query searched, int32 index table materialised); Chamfer = two ngpd_nn_sqdist_reduce
passes (block reduction fused, no per-point output) between the noisy cloud and its clean surface, indices built beforehand.
Several GPUs: every rank generates its chunks; kNN = partition.ShardedKnn (Morton slabs planned from the shards, one index per
rank over slab + halo, the halo width checked against the k-th distances); Chamfer = partition.sharded_chamfer (replicated
targets, sharded queries, fused reduction, one all-reduce).  Markdown table on stdout (rank 0)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import ngpd_b200
from ngpd_b200 import _lib, partition, workloads as W

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nmin = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, reps=3):
    # two warm calls: the result tensor of call i is still alive while call i + 1 allocates its own, so the caching allocator needs
    # two blocks before the timed region stops calling cudaMalloc (a 640 MB cudaMalloc costs 1-10 ms, more on some boxes)
    fn(); fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(line):
    if rank == 0:
        print(line, flush=True)


say(f"| points | GPUs | k | index build ms | kNN ms | G queries/s | algorithmic GB/s | frac of {world} x {peak:.0f} GB/s | halo need / width | Chamfer (both directions) ms | G points/s | CD |")
say("|---|---|---|---|---|---|---|---|---|---|---|---|")
for n in (1_000_000, 10_000_000, 100_000_000, 1_000_000_000):
    if n > nmax or n < nmin:
        continue
    sigma = 0.3 * W.mean_knn_distance("creased", n, 6)
    cs, ns, gs = [], [], []
    for c in W.chunks_of(n, rank, world):
        p, _, g = W.surface_chunk("creased", n, c, 1234, dev)
        cs.append(p); ns.append(W.noise_chunk(p, sigma, c)); gs.append(g)
    clean, noisy, gids = torch.cat(cs), torch.cat(ns), torch.cat(gs)
    del cs, ns, gs
    ks = (8, 16, 32, 64) if n * 64 * 4 / world < 60e9 else ((8, 16, 32) if n * 32 * 4 / world < 60e9 else (8, 16))
    for k in ks:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        halo = "-"
        if world == 1:
            grid = _lib.Grid(noisy, k)
            torch.cuda.synchronize(); build = (time.perf_counter() - t0) * 1e3
            out = {}
            ms = timed(lambda: out.__setitem__("t", grid.knn(noisy, k, _lib.KNN_QUERY_IS_TREE)))
            out.clear()
        else:
            sk = partition.ShardedKnn(noisy, gids, k)
            torch.cuda.synchronize(); dist.barrier(); build = (time.perf_counter() - t0) * 1e3
            out = {}
            ms = timed(lambda: out.__setitem__("t", sk.grid.knn(sk.queries, k, 0)))
            out.clear()
            if n * k * 8 / world < 70e9:
                sk.knn(k, global_ids=False)
                halo = f"{sk.check():.4g} / {sk.plan.halo_width:.4g}"
            grid = None
        cham = "- | - | -"
        if k == 16:
            if world == 1:
                gc = _lib.Grid(clean, 16)
                res = {}
                cms = timed(lambda: res.__setitem__("v", (grid.nn_reduce(clean), gc.nn_reduce(noisy))))
                a, b = res["v"]
                cd = float((a[0] + b[0]) / (a[3] + b[3]))
                del gc
            else:
                del sk
                torch.cuda.empty_cache()
                res = {}
                cms = timed(lambda: res.__setitem__("v", partition.sharded_chamfer(clean, noisy)), reps=1)
                cd = res["v"]["chamfer"]
            cham = f"{cms:.2f} | {2 * n / (cms * 1e-3) / 1e9:.2f} | {cd:.6e}"
        gbs = (24 + 4 * k) * n / (ms * 1e-3) / 1e9
        say(f"| {n} | {world} | {k} | {build:.1f} | {ms:.2f} | {n / (ms * 1e-3) / 1e9:.3f} | {gbs:.0f} | {gbs / (peak * world):.3f} | {halo} | {cham} |")
        grid = sk = None
        torch.cuda.empty_cache()
    del clean, noisy, gids
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
