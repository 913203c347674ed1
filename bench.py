#!/usr/bin/env python
"""Benchmark of the denoising hot path: point-iterations/s of one Processor.denoise iteration body
(kNN(k_f) + NVT + smoothing + NVT + labels + class-wise update) on the synthetic surfaces of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl ours|reference]
                    [--surface creased|cubes] [--strategy flat/edge/feature] [--k-feature 16] [--clamp]

Prints ONE JSON line.  Keys beyond the driver's contract:
  cold            iterations 1 and 2 of a fresh session (what one Processor.denoise() call costs; `value` is a warm step)
  validated       sampled rows of the full-size run re-answered by the exact shell search, by an fp64 brute force in torch and
                  by the oracle (labels, smoothed normals) -- the benchmark proves what it ran
  checksum        order- and partition-independent digest of the final state: equal across --gpus 1/2/4/8 iff the slab runs
                  reproduce the single-GPU run bit for bit
  class_histogram labels of the last iteration
`--impl reference` times the CPU oracle port of the reference's implementation on the host cores (bounded sample)."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_F, K_U = 16, 8
ALPHAS = (1.0, 0.2, 1.0)
METRIC = "denoise point-iterations/sec (kNN+NVT+update)"


def algorithmic_bytes(k_f=None, k_u=None):
    """compulsory bytes per point per launch of each kernel (fp32 x3 vectors, int32 indices)."""
    k_f, k_u = K_F if k_f is None else k_f, K_U if k_u is None else k_u
    return {
        "knn": 12 + 12 + 4 * k_f,                    # query, tree point, neighbour row out
        "nvt_smooth": 4 * k_f + 12 + 12 + 12,        # row in, position, normal, smoothed normal out
        "nvt_classify": 4 * k_f + 12 + 12 + 1 + 12,  # row in, position, smoothed normal, label + crease vector out
        "flat_scalars": None,                        # SURVEY 8(d) counts two passes over the class' rows (2 x (4 k_u + 1) B); the sums ride in the
                                                     # stage-2 kernel and delta comes from its per-block maxima: no pass over the rows is left to rate
        "update": 4 * k_u + 12 + 12 + 1 + 12 + 12,   # phase C of 8(d), all three class launches together
        "iteration": 134 + 8 * k_f + 4 * k_u,
    }


def ncu_traffic(n):
    """DRAM bytes per launch of the kernels of one group, from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per launch and point at the captured size), scaled to n points."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}, None
    t = json.load(open(p))
    per = {k: sum(v) / len(v) for k, v in t["bytes_per_point"].items()}
    groups = t.get("groups") or {
        "knn": ["session_knn_rerank_kernel", "session_knn_fast_kernel", "session_knn_wide_kernel", "session_knn_fix_kernel"],
        "nvt_smooth": ["session_nvt_smooth_kernel", "session_nvt_smooth_late_kernel"], "nvt_classify": ["session_nvt_classify_kernel"],
        "flat_scalars": ["session_partial_reduce_kernel", "session_center_kernel", "session_class_max_kernel"]}
    out = {g: sum(per.get(k, 0.0) for k in ks) * n for g, ks in groups.items()}
    if "update" not in out:
        upd = t["bytes_per_point"]
        out["update"] = (sum(upd.get("session_update_kernel", [0.0])) + sum(upd.get("session_update_rows_kernel", [0.0])) +
                         sum(upd.get("session_apply_rows_kernel", [0.0]))) * n
    return out, f"{t['source']} at {t['points']} points, scaled per point"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# input: generated chunk by chunk, every rank only the chunks of its share (no rank ever builds the whole cloud when N > 1)
# ---------------------------------------------------------------------------------------------------------------------
def make_shard(args, n, device, rank, world):
    """(noisy positions, analytic normals, global ids) of this rank's chunks.  Noise: isotropic Gaussian of sigma = 0.3 x the
    expected mean 6-NN edge length of the surface (analytic, so that no pass over the whole cloud is needed)."""
    import torch
    from ngpd_b200 import workloads as W
    sigma = 0.3 * W.mean_knn_distance(args.surface, n, 6)
    ps, ns, gs = [], [], []
    for c in W.chunks_of(n, rank, world):
        p, q, g = W.surface_chunk(args.surface, n, c, 1234, device)
        ps.append(W.noise_chunk(p, sigma, c)); ns.append(q); gs.append(g)
    if not ps:
        z = torch.empty((0, 3), device=device)
        return z, z.clone(), torch.empty(0, dtype=torch.long, device=device)
    return torch.cat(ps), torch.cat(ns), torch.cat(gs)


def single_gpu_normals(noisy, analytic):
    """PCA normals of the noisy cloud over its own 12-NN graph (GraphBuilder.py:60-63, 95-111), oriented by the analytic
    normal (stand-in for the reference's O(N^2) spanning-tree orientation, which is preprocessing)."""
    import torch
    from ngpd_b200 import _lib
    n = noisy.size(0)
    grid = _lib.Grid(noisy, 12)
    table = grid.knn(noisy, 12, _lib.KNN_SKIP_SELF | _lib.KNN_QUERY_IS_TREE)
    nrm = torch.empty_like(noisy)
    _lib.check(_lib.load().ngpd_pca_normals(noisy.data_ptr(), table.data_ptr(), None, n, 12, nrm.data_ptr(), None, None, _lib.stream()), "pca")
    flip = (nrm * analytic).sum(1) < 0
    nrm[flip] *= -1
    del grid, table
    torch.cuda.synchronize()
    return nrm


def strategy_of(args):
    from ngpd_b200 import _lib
    names = {"flat": _lib.STEP_FLAT, "edge": _lib.STEP_EDGE, "feature": _lib.STEP_FEATURE, "corner": _lib.STEP_CORNER, "none": _lib.STEP_NONE}
    parts = args.strategy.split("/")
    assert len(parts) == 3 and all(p in names for p in parts), args.strategy
    return tuple(names[p] for p in parts)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's iteration body (test infrastructure; used here only as the timed baseline)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_iteration_rate(args, n_sample, iters, workers=-1, warmup=1):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import ngpd_oracle as O
    from ngpd_b200 import workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores if workers == -1 else 1)
    sigma = 0.3 * W.mean_knn_distance(args.surface, n_sample, 6)
    ps, ns = [], []
    for c in W.chunks_of(n_sample):
        p, q, _ = W.surface_chunk(args.surface, n_sample, c, 4321, "cpu")
        ps.append(W.noise_chunk(p, sigma, c)); ns.append(q)
    noisy, nrm = torch.cat(ps).numpy(), torch.cat(ns).numpy()
    knn = lambda t, q, k: O.knn_kdtree(t, q, k, workers=workers)
    xt = O.acos_threshold(math.pi * 5 / 12)
    d = np.float32(2) * O.average_edge_length(noisy, knn(noisy, noisy, 6))
    pos = noisy
    strat = tuple(args.strategy.split("/"))
    for _ in range(warmup):
        pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, strategy=strat, knn=knn)
    t0 = time.perf_counter()
    for _ in range(iters):
        pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, strategy=strat, knn=knn)
    dt = time.perf_counter() - t0
    torch.set_num_threads(cores)
    return n_sample * iters / dt, (cores if workers == -1 else 1), dt / iters


def run_reference(args):
    """--impl reference: the CPU implementation (oracle port; the reference itself is Python and cannot travel)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_sample = args.ref_points
    value, cores, per_it = cpu_iteration_rate(args, n_sample, args.steps, -1, args.warmup)
    single, _, per_it1 = cpu_iteration_rate(args, min(n_sample, 100_000), 2, 1, 1)
    sample = (f"{n_sample}-point cloud from the same generator (the GPU arm's workload has {args.points} points; throughput per point is "
              f"size-independent to first order for a KD-tree pipeline), {args.steps} timed iterations of the oracle port "
              f"(NumPy + SciPy cKDTree workers=-1 + LAPACK via torch, all {cores} host threads).  As the reference calls it "
              f"(KDTree.query single-threaded, one torch thread; 100 k points): {single:.0f} point-iterations/s")
    cfg = workload_config(args, n_sample)
    cfg["full_points"] = args.points
    cfg["note"] = "points = the sample this arm ran; full_points = the GPU arm's cloud"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "point-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_it * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "point-iterations/s", "cores": cores, "kind": "port", "sample": sample,
                             "single_thread_value": single},
            "e2e": {"value": value, "unit": "point-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n):
    if args.surface == "cubes":
        name = (f"synthetic lattice of 17^3 small cubes (dense creases: about a fifth of the points sit on an edge or a corner), {n} points, "
                f"Gaussian noise sigma = 0.3 x mean 6-NN distance, k={K_F}/{K_U}; one Processor.denoise iteration body per step")
    elif K_F == 32:
        name = ("BASELINE.json configs[2] stand-in (xyzrgb_dragon is a missing blob, SURVEY 8d): the synthetic creased surface at "
                f"{n} points, Gaussian noise, k={K_F}/{K_U}, one Processor.denoise iteration body per step")
    else:
        name = ("BASELINE.json configs[3]: synthetic 100M-point noisy CAD-like surface (cube faces 60 % + torus 40 %, isotropic Gaussian noise "
                f"sigma = 0.3 x mean 6-NN distance), k={K_F}/{K_U}; one Processor.denoise iteration body per step; the whole cloud "
                "fits one B200, more GPUs split the same cloud into Morton slabs")
    return {"workload": name,
            "points": n, "k_feature": K_F, "k_update": K_U, "strategy": args.strategy, "alpha": list(ALPHAS), "surface": args.surface,
            "clamp_to_original": bool(args.clamp),
            "l2": "inputs larger than L2 (positions+normals+neighbour table >> 126 MB); no flush needed",
            "partition": "single GPU" if args.gpus == 1 else
                         f"{args.gpus} Morton slabs + halo exchange; input generated per rank, slabs planned from the shards (no rank holds the cloud)"}


# ---------------------------------------------------------------------------------------------------------------------
# validation of the full-size run (sampled rows)
# ---------------------------------------------------------------------------------------------------------------------
def validate(args, sess, params, tree_local, n_owned, device, sample_rows=10000, brute_rows=256, refresh=None):
    """Runs the feature phases of one more iteration (state unchanged: no update, normals not committed) and checks sampled
    owned rows: (1) neighbour rows == the exact fp64 shell search of the public ngpd_knn (NGPD_KNN_EXACT_ONLY) over the same
    tree; (2) a subset == an fp64 brute force over ALL tree points written in torch (no kernel of this repo involved);
    (3) smoothed normals and labels == the NumPy oracle, teacher-forced on the session's neighbour rows / smoothed normals."""
    import ctypes
    import numpy as np
    import torch
    from ngpd_b200 import _lib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ngpd_oracle as O
    lib = _lib.load()
    n_local = tree_local.size(0)
    g = torch.Generator(device="cpu"); g.manual_seed(2024)
    m = min(sample_rows, n_owned)
    sample = torch.randperm(n_owned, generator=g)[:m].to(device)
    pos0, nrm0, _ = sess.get_state(False)                       # original (local) order
    ref = ctypes.byref(params)
    _lib.check(lib.ngpd_session_phase_features(sess._h, ref, 0, _lib.stream()), "features 0")
    if refresh is not None:
        refresh(2)                                               # slabs: the halo copies of the smoothed normals
    fn_view = torch.as_tensor(_DevView(lib.ngpd_session_buffer(sess._h, 2), (n_local, 4), "<f4"), device=device)
    fn_tree = fn_view.clone()                                    # smoothed normals, tree order (halo rows of a slab: as last refreshed)
    _lib.check(lib.ngpd_session_phase_features(sess._h, ref, 1, _lib.stream()), "features 1")
    idx_view = torch.as_tensor(_DevView(lib.ngpd_session_buffer(sess._h, 6), (n_local, params.k_feature), "<i4"), device=device)
    lab_view = torch.as_tensor(_DevView(lib.ngpd_session_buffer(sess._h, 5), (n_local,), "|u1"), device=device)
    perm = sess.order().long()                                   # tree position -> original (local) index
    inv = torch.empty(n_local, dtype=torch.int32, device=device)
    inv[perm] = torch.arange(n_local, dtype=torch.int32, device=device)
    srow = inv[sample].long()
    rows_session = perm[idx_view[srow].long()]                   # [m, k] original (local) indices
    lab_session = lab_view[srow].cpu().numpy()
    fn_session = fn_tree[srow, :3].cpu().numpy()
    q = pos0[sample].contiguous()
    # (1) exact shell search through the public ABI over a separately built index
    grid = _lib.Grid(tree_local, params.k_feature)
    rows_exact = grid.knn(q, params.k_feature, _lib.KNN_EXACT_ONLY).long()
    del grid
    eq_exact = int((rows_exact == rows_session).all(dim=1).sum())
    # (2) fp64 brute force in torch: ((dx^2 + dy^2) + dz^2), ties by index (none expected in random data)
    b = min(brute_rows, m)
    eq_brute = 0
    for i in range(b):
        qd = q[i].double()
        d2 = (tree_local[:, 0].double() - qd[0]) ** 2
        d2 += (tree_local[:, 1].double() - qd[1]) ** 2
        d2 += (tree_local[:, 2].double() - qd[2]) ** 2
        near = torch.topk(d2, params.k_feature, largest=False, sorted=True).indices
        eq_brute += int(torch.equal(near, rows_session[i]))
        del d2
    # (3) oracle on the compacted neighbourhoods
    uniq, remap = torch.unique(torch.cat([rows_session.reshape(-1), sample]), return_inverse=True)
    nbr_c = remap[:m * params.k_feature].reshape(m, params.k_feature).cpu().numpy()
    rows_c = remap[m * params.k_feature:].cpu().numpy()
    pos_c, nrm_c = pos0[uniq].cpu().numpy(), nrm0[uniq].cpu().numpy()
    fn_c = fn_tree[inv[uniq].long(), :3].cpu().numpy()
    xt = np.float32(params.x_thresh)
    w1, V1, _, _ = O.nvt(pos_c, nrm_c, rows_c, nbr_c, xt)
    fn_o = O.smooth_normals(w1, V1, nrm_c[rows_c], params.tau, params.damp)
    ang = np.arctan2(np.linalg.norm(np.cross(fn_o.astype(np.float64), fn_session.astype(np.float64)), axis=1),
                     (fn_o.astype(np.float64) * fn_session).sum(1))
    w2, _, _, _ = O.nvt(pos_c, fn_c, rows_c, nbr_c, xt)
    lab_o = O.classes(w2, params.scale)
    out = {"rows_sampled": m, "knn_rows_equal_exact_search": eq_exact, "knn_rows_bruteforce_fp64": b, "knn_rows_equal_bruteforce": eq_brute,
           "labels_equal_oracle": int((lab_o == lab_session).sum()), "smoothed_normals_within_1e-4_rad_of_oracle": int((ang <= 1e-4).sum()),
           "smoothed_normals_max_angle": float(ang.max())}
    del pos0, nrm0, fn_tree, perm, inv
    torch.cuda.empty_cache()
    return out


def run_extra(args, name, n, dev, surface, k_f, k_u, strategy, clamp, steps=6, warmup=4):
    """A secondary configuration on one GPU (same generator, same timing rules), so that the driver's default run also shows
    BASELINE configs[2] (k = 32 at 10 M points) and a surface that exercises edge_step / feature_step."""
    import argparse
    import torch
    from ngpd_b200 import _lib
    a = argparse.Namespace(**{**vars(args), "surface": surface, "strategy": strategy, "clamp": clamp})
    noisy, analytic, _ = make_shard(a, n, dev, 0, 1)
    nrm = single_gpu_normals(noisy, analytic)
    del analytic
    sess = _lib.Session(noisy, k_f)
    sess.reserve(k_f)
    sess.set_state(noisy, nrm)
    s, c = sess.mean_edge_length_parts(6)
    d = 2.0 * s / c
    params = _lib.make_params(k_f, k_u, None, 0.3, 3.0, 0.2, strategy_of(a), ALPHAS, d * (20000.0 if clamp else 1.0),
                              _lib.STEP_SNAPSHOT_CLASSES if clamp else 0, d if clamp else 0.0)
    if clamp:
        sess.set_original(noisy)
    for _ in range(warmup):
        sess.step(params)
    sess.set_profiling(True)
    sess.get_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        sess.step(params)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = sess.get_profile()
    digest = sess.checksum()
    tot = max(digest[6] + digest[7] + digest[8], 1)
    peak, _ = measured_peak()
    b_iter = 134 + 8 * k_f + 4 * k_u
    out = {"name": name, "points": n, "surface": surface, "k_feature": k_f, "k_update": k_u, "strategy": strategy, "clamp_to_original": clamp,
           "steps": steps, "warmup": warmup, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "point-iterations/s",
           "iteration_frac_of_hbm_peak": b_iter * n / (ms * 1e-3) / 1e9 / peak,
           "class_histogram": {"flat": digest[6], "edge": digest[7], "corner": digest[8], "edge_fraction": digest[7] / tot,
                               "corner_fraction": digest[8] / tot},
           "kernels_ms_per_step": {k: v[0] / steps for k, v in prof.items() if v[1]}}
    del sess, noisy, nrm
    torch.cuda.empty_cache()
    return out


class _DevView:
    """zero-copy torch view of a raw device pointer"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": shape, "typestr": typestr, "version": 2}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ngpd_b200
    from ngpd_b200 import _lib, partition
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.points
    peak, peak_src = measured_peak()
    bytes_per = algorithmic_bytes()
    strategy = strategy_of(args)
    flags = _lib.STEP_SNAPSHOT_CLASSES if args.clamp else 0

    noisy, analytic, gids = make_shard(args, n, dev, rank, world)
    if world == 1:
        nrm = single_gpu_normals(noisy, analytic)
        del analytic, gids
        sess = _lib.Session(noisy, K_F)
        sess.reserve(K_F)
        sess.set_state(noisy, nrm)
        s, c = sess.mean_edge_length_parts(6)
        d = 2.0 * s / c
        params = _lib.make_params(K_F, K_U, None, 0.3, 3.0, 0.2, strategy, ALPHAS, d * (20000.0 if args.clamp else 1.0), flags,
                                  d if args.clamp else 0.0)
        if args.clamp:
            sess.set_original(noisy)
        step = lambda: sess.step(params)
        profile_src, tree_local, n_local = sess, noisy, n
        slab = None
    else:
        slab = partition.SlabSession(noisy, analytic, K_F, K_U, ALPHAS, strategy=strategy, shard_ids=gids, flags=flags)
        slab.session.reserve(K_F)
        slab.pca_normals(12, orient_like="current")
        balance = None
        if not args.no_balance:
            # cost-aware slab sizes (set-up, untimed): a trial on a throw-away session (2 + 3 iterations, the last 3 measured) gives every
            # slab's kernel time, the slabs are cut again so that they cost the same, and the run starts over from the same input
            t = slab.rank_costs(3, 2)
            fr = partition.rebalance_fractions(t, slab.plan.fractions)
            balance = {"trial_kernel_ms_by_rank": [round(x, 4) for x in t], "point_fractions": [round(x, 5) for x in fr]}
            del slab
            torch.cuda.empty_cache()
            slab = partition.SlabSession(noisy, analytic, K_F, K_U, ALPHAS, strategy=strategy, shard_ids=gids, flags=flags, fractions=fr)
            slab.session.reserve(K_F)
            slab.pca_normals(12, orient_like="current")
        del noisy, analytic, gids
        if args.clamp:
            d = 2.0 * slab.mean_edge_length
            slab.params.dmax, slab.params.clamp_radius = d * 20000.0, d
            slab.session.set_original(slab.tree_local)
        sess, params = slab.session, slab.params
        step = slab.step
        profile_src, tree_local, n_local = slab.session, slab.tree_local, slab.n_owned
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    def timed(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- cold: iterations 1 and 2 of a fresh session (no stored candidates: the first search walks the grid for every row)
    cold1 = timed(1)
    cold2 = timed(1)
    cold = {"ms_iteration_1": cold1, "ms_iteration_2": cold2, "value": n * 2 / ((cold1 + cold2) * 1e-3), "unit": "point-iterations/s",
            "note": "Processor.denoise() is two iterations from a fresh index (Processor.py:123); `value` of this line is a steady-state iteration"}
    for _ in range(max(args.warmup - 2, 0)):
        step()
    profile_src.set_profiling(True)
    profile_src.get_profile()
    launches_before = profile_src.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(args.steps)
    prof = profile_src.get_profile()
    profile_src.set_profiling(False)
    by_rank = None
    if world > 1:
        # every rank's time per kernel group: the spread between the ranks is what the cross-rank rounds ("halo") wait for
        mine = torch.tensor([prof[k][0] / args.steps for k in _lib.Session.PROFILE_NAMES], dtype=torch.float64, device=dev)
        table = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(table, mine)
        by_rank = {k: [round(float(t[i]), 4) for t in table] for i, k in enumerate(_lib.Session.PROFILE_NAMES)}
    # single GPU: the fused step restarts its counter, so the last step's count x steps; slabs likewise (one C call per step)
    launches = profile_src.launch_count() * args.steps
    clocks = sampler.stop() if sampler else None
    value = n * args.steps / (ms * 1e-3)
    t0, t1, t2 = sess.knn_stats()                               # rows of the last search that each tier handed on (rank 0's slab)
    knn_tiers = {"rows": n_local, "reranked_from_stored_candidates": n_local - t0, "searched_3x3x3": t0 - t1, "searched_5x5x5": t1 - t2,
                 "exact_shell_search": t2}

    # ---- what the run computed: digest of the final state (compare across --gpus), labels of the last iteration, halo margin
    if world == 1:
        digest = sess.checksum()
        halo = None
    else:
        digest = slab.checksum()
        need = slab.verify_halo()                               # raises if a slab search could have missed a foreign point
        halo = {"halo_width": slab.plan.halo_width, "needed": need, "rows_owned_rank0": slab.n_owned, "rows_halo_rank0": slab.n_halo,
                "balance": balance}
    names = _lib.Session.CHECKSUM_FIELDS
    checksum = {k: (f"{v:016x}" if k.endswith("hash") else (v - (1 << 64) if v >= (1 << 63) else v)) for k, v in zip(names, digest)}
    checksum["after_iterations"] = 2 + max(args.warmup - 2, 0) + args.steps
    hist = {"flat": digest[6], "edge": digest[7], "corner": digest[8]}
    tot = max(sum(hist.values()), 1)
    class_histogram = {**hist, "edge_fraction": hist["edge"] / tot, "corner_fraction": hist["corner"] / tot}

    validated = None
    if not args.no_validate:
        per_rank = max(10000 // world, 1250)
        v = validate(args, sess, params, tree_local, n_local, dev, per_rank, max(256 // world, 32), slab._refresh if slab else None)
        if world > 1:
            keys = [k for k in v if k != "smoothed_normals_max_angle"]
            t = torch.tensor([v[k] for k in keys], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            a = torch.tensor([v["smoothed_normals_max_angle"]], dtype=torch.float64, device=dev)
            dist.all_reduce(a, op=dist.ReduceOp.MAX)
            v = {**{k: int(x) for k, x in zip(keys, t.tolist())}, "smoothed_normals_max_angle": float(a.item())}
        # neighbour rows: bit-exact.  Labels depend on eigenvalues only (well conditioned): all but rounding cases.  Smoothed normals
        # depend on LAPACK's eigenvector signs: the oracle's own MKL agrees with any other solver on ~99.8 % of rows.
        v["ok"] = bool(v["knn_rows_equal_exact_search"] == v["rows_sampled"] and v["knn_rows_equal_bruteforce"] == v["knn_rows_bruteforce_fp64"]
                       and v["labels_equal_oracle"] >= 0.999 * v["rows_sampled"]
                       and v["smoothed_normals_within_1e-4_rad_of_oracle"] >= 0.99 * v["rows_sampled"])
        v["how"] = ("after the timed loop, at the full size: rows of a seeded sample of the owned points vs ngpd_knn(NGPD_KNN_EXACT_ONLY) over a "
                    "separately built index, vs an fp64 brute force in torch, and labels / smoothed normals vs the NumPy oracle "
                    "(teacher-forced on the session's rows)")
        validated = v

    # ---- end to end through the host-buffer entry point: pinned host -> device, one iteration, device -> host
    e2e = None
    if world == 1:
        with partition.near_gpu(dev.index) as cpus:
            pos_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()
            nrm_h = torch.empty((n, 3), dtype=torch.float32).pin_memory()
            pos_o = torch.empty((n, 3), dtype=torch.float32).pin_memory()
            nrm_o = torch.empty((n, 3), dtype=torch.float32).pin_memory()
            lab_o = torch.empty(n, dtype=torch.uint8).pin_memory()
        p, q, _ = sess.get_state(False)
        pos_h.copy_(p); nrm_h.copy_(q)
        del p, q
        e_steps = max(2, min(args.steps, 5))
        sess.run_host(params, 1, pos_h, nrm_h, pos_o, nrm_o, lab_o)          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            sess.run_host(params, 1, pos_h, nrm_h, pos_o, nrm_o, lab_o)
            pos_h, pos_o = pos_o, pos_h
            nrm_h, nrm_o = nrm_o, nrm_h
        dt = time.perf_counter() - t0
        e2e = {"value": n * e_steps / dt, "unit": "point-iterations/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 25,
               "steps": e_steps, "call": "ngpd_session_run_host (pinned host buffers in and out, frozen index resident)"}
        del pos_h, nrm_h, pos_o, nrm_o, lab_o
    else:
        # every rank streams its own slab through pinned host buffers; halo rows come from their owners over NVLink
        k = slab.n_owned
        _, p, q, _ = slab.owned_state()
        pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
        with partition.near_gpu(dev.index) as cpus:
            pos_h, nrm_h, pos_o, nrm_o, lab_o = pin(k, 3), pin(k, 3), pin(k, 3), pin(k, 3), pin(k, dtype=torch.uint8)
        pos_h.copy_(p); nrm_h.copy_(q)
        del p, q
        e_steps = max(2, min(args.steps, 5))
        slab.step_host(pos_h, nrm_h, pos_o, nrm_o, lab_o)                        # warm-up
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            slab.step_host(pos_h, nrm_h, pos_o, nrm_o, lab_o)
            pos_h, pos_o = pos_o, pos_h
            nrm_h, nrm_o = nrm_o, nrm_h
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": n * e_steps / float(t.item()), "unit": "point-iterations/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 25,
               "steps": e_steps, "call": "SlabSession.step_host (each rank: pinned host buffers of its slab in and out, halo rows by peer stores over NVLink)"}

    # ---- the metric's second half: kNN queries/s through the public ngpd_knn (no temporal coherence: every query searched)
    knn_line = None
    if world == 1 and not args.no_knn:
        del sess
        profile_src = step = None
        torch.cuda.empty_cache()
        grid = _lib.Grid(tree_local, K_F)
        fl = _lib.KNN_QUERY_IS_TREE
        for _ in range(2):
            tab = grid.knn(tree_local, K_F, fl)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(3):
            tab = grid.knn(tree_local, K_F, fl)
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / 3
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2 = grid.nn_sqdist(tree_local, False, fl)
        c0.record()
        for _ in range(3):
            d2 = grid.nn_sqdist(tree_local, False, fl)
        c1.record()
        torch.cuda.synchronize()
        cms = c0.elapsed_time(c1) / 3
        knn_line = {"metric": "kNN queries/sec", "k": K_F, "value": n / (kms * 1e-3), "ms": kms, "algorithmic_bytes_per_query": 24 + 4 * K_F,
                    "frac_of_hbm_peak": (24 + 4 * K_F) * n / (kms * 1e-3) / 1e9 / peak,
                    "nearest_neighbour_queries_per_s": n / (cms * 1e-3), "nearest_ms": cms,
                    "call": "ngpd_knn / ngpd_nn_sqdist (public ABI, queries = the tree's own noisy points, index output materialised)"}
        del tab, d2, grid

    extra = None
    if world == 1 and not args.no_extra:
        del tree_local
        torch.cuda.empty_cache()
        m = min(n, 10_000_000)
        extra = [run_extra(args, "BASELINE configs[2] stand-in: k = 32 / 8 at 10 M points", m, dev, "creased", 32, 8, "flat/edge/feature", False),
                 run_extra(args, "dense creases: lattice of 17^3 cubes, Processor.denoise strategy", m, dev, "cubes", 16, 8, "flat/edge/feature", False),
                 run_extra(args, "dense creases, the notebook's CTD-QEM row (feature_step on every point)", m, dev, "cubes", 16, 8,
                           "feature/feature/feature", False),
                 run_extra(args, "dense creases, the notebook's 'Ours' row (flat/feature/feature from one snapshot + displacement clamp)", m, dev,
                           "cubes", 16, 8, "flat/feature/feature", True)]
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    kernels = {}
    step_ms = ms / args.steps
    for name, (tms, cnt) in prof.items():
        if cnt == 0:
            continue
        per_step = tms / args.steps
        entry = {"ms_per_step": per_step, "launches_per_step": cnt / args.steps, "share_of_step": per_step / step_ms}
        if bytes_per.get(name) is not None:
            ab = bytes_per[name] * n_local
            entry.update({"algorithmic_bytes_per_point": bytes_per[name], "achieved_gbs": ab / (per_step * 1e-3) / 1e9,
                          "frac": ab / (per_step * 1e-3) / 1e9 / peak})
        elif name == "halo":
            entry["note"] = "halo refreshes + cross-rank scalars; includes waiting for the slowest peer"
        else:
            entry["note"] = "flat_step's centre / delta: reduction of per-block sums and maxima left by the stage-2 kernel + the candidate blocks"
        kernels[name] = entry
    kernel_ms = sum(v["ms_per_step"] for v in kernels.values())
    dom = max((k for k in kernels if "frac" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    traffic, traffic_src = ncu_traffic(n_local)
    for name in kernels:
        kernels[name]["ncu_dram_bytes_per_step"] = traffic.get(name)
    it_gbs = bytes_per["iteration"] * n / (step_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "per GPU (rank 0's rows over rank 0's kernel time); kNN is instruction/latency-bound, reported against HBM as SURVEY 8(d) asks",
                "iteration_achieved_gbs": it_gbs, "iteration_achieved_gbs_per_gpu": it_gbs / world,
                "iteration_frac": it_gbs / (world * peak),
                "outside_kernels_ms_per_step": step_ms - kernel_ms}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, per_it = cpu_iteration_rate(args, args.cpu_points, 5)          # ~11 s of CPU work at 200 k points
        cpu = {"value": v, "unit": "point-iterations/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_points}-point cloud from the same generator, 1 warm-up + 5 timed iterations of the oracle port "
                         f"(NumPy + SciPy KD-tree workers=-1 + LAPACK), {per_it:.2f} s per iteration"}
    line = {"metric": METRIC, "value": value, "unit": "point-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n), "clocks": clocks,
            "e2e": e2e, "cold": cold, "validated": validated, "checksum": checksum, "class_histogram": class_histogram, "halo": halo,
            "knn": knn_line, "knn_tiers_last_step": knn_tiers, "kernel_ms_per_step_by_rank": by_rank, "extra_configs": extra, "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global K_F, K_U
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=int(os.environ.get("NGPD_BENCH_POINTS", 100_000_000)))
    ap.add_argument("--cpu-points", type=int, default=200_000)
    ap.add_argument("--ref-points", type=int, default=400_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-validate", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: equal-size slabs instead of slabs of equal measured cost")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary single-GPU configurations (k = 32; dense-crease surface)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k-feature", type=int, default=K_F, help="k of the feature pass (configs[2]: 32)")
    ap.add_argument("--k-update", type=int, default=K_U)
    ap.add_argument("--surface", default="creased", choices=["creased", "cubes"],
                    help="creased: cube + torus of configs[3] (~1 %% crease points); cubes: lattice of small cubes (~20 %% crease / corner points)")
    ap.add_argument("--strategy", default="flat/edge/feature", help="step per class 0/1/2: flat, edge, feature, corner or none "
                    "(feature/feature/feature = the notebook's CTD-QEM row, PostProcessing.ipynb#c9)")
    ap.add_argument("--clamp", action="store_true", help="the notebook's 'Ours' row: every class from one snapshot, d x 20000 inside the steps, "
                    "then |x - x_original| < d (PostProcessing.ipynb#c9)")
    args = ap.parse_args()
    K_F, K_U = args.k_feature, args.k_update
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as ge
        ge.build()                                              # every rank; serialised by a file lock, the library is replaced atomically
        run_ours(args)


if __name__ == "__main__":
    main()
