#!/usr/bin/env python
"""Benchmark of the denoising hot path: point-iterations/s of one Processor.denoise iteration body
(kNN(k_f) + NVT + smoothing + NVT + labels + class-wise update) on the synthetic creased surface of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--impl ours|reference]

Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU oracle port of the reference's
implementation on the host cores (bounded sample), for the driver's own speed-up ratio."""
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
sys.path.insert(0, os.path.join(ROOT, "oracle"))

K_F, K_U = 16, 8
ALPHAS = (1.0, 0.2, 1.0)


def algorithmic_bytes(k_f=None, k_u=None):
    """compulsory bytes per point per launch of each kernel (fp32 x3 vectors, int32 indices)."""
    k_f, k_u = K_F if k_f is None else k_f, K_U if k_u is None else k_u
    return {
        "knn": 12 + 12 + 4 * k_f,                    # query, tree point, neighbour row out
        "nvt_smooth": 4 * k_f + 12 + 12 + 12,        # row in, position, normal, smoothed normal out
        "nvt_classify": 4 * k_f + 12 + 12 + 1 + 12,  # row in, position, smoothed normal, label + crease vector out
        "flat_scalars": 2 * (4 * k_u + 1),           # two passes over the class' rows (gathers are cache hits)
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
    groups = {"knn": ("session_knn_rerank_kernel", "session_knn_fast_kernel", "session_knn_wide_kernel", "session_knn_fix_kernel"),
              "nvt_smooth": ("session_nvt_smooth_kernel", "session_nvt_smooth_late_kernel"), "nvt_classify": ("session_nvt_classify_kernel",),
              "flat_scalars": ("session_partial_reduce_kernel", "session_center_kernel", "session_class_max_kernel")}
    out = {g: sum(per.get(k, 0.0) for k in ks) * n for g, ks in groups.items()}
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


def make_input(n, device, seed=1234):
    """clean surface -> noisy positions (sigma = 0.3 x mean 6-NN distance, random direction) + PCA normals oriented by the
    analytic normal (stand-in for the reference's O(N^2) spanning-tree orientation, which is preprocessing)."""
    import torch
    from ngpd_b200 import _lib, workloads
    clean, normal = workloads.creased_surface(n, seed, device)
    sess = _lib.Session(clean, 16)
    s, c = sess.mean_edge_length_parts(6)
    l6 = s / c * 6.0 / 5.0                                   # the 6-NN row holds the zero self edge
    del sess
    noisy = workloads.add_noise(clean, 0.3 * l6)
    grid = _lib.Grid(noisy, 12)
    table = grid.knn(noisy, 12, _lib.KNN_SKIP_SELF | _lib.KNN_QUERY_IS_TREE)
    nrm = torch.empty_like(noisy)
    _lib.check(_lib.load().ngpd_pca_normals(noisy.data_ptr(), table.data_ptr(), None, n, 12, nrm.data_ptr(), None, None, _lib.stream()), "pca")
    flip = (nrm * normal).sum(1) < 0
    nrm[flip] *= -1
    del grid, table, clean, normal
    torch.cuda.synchronize()
    return noisy, nrm


def cpu_iteration_rate(n_sample, iters, seed=4321):
    """the oracle port of the reference's iteration body (NumPy + SciPy KD-tree + LAPACK via torch) on the host cores"""
    import numpy as np
    import torch
    import ngpd_oracle as O
    from ngpd_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clean, normal = workloads.creased_surface(n_sample, seed, "cpu")
    sigma = 0.3 * workloads.expected_spacing(n_sample)
    noisy = workloads.add_noise(clean, sigma).numpy()
    nrm = normal.numpy()
    knn = lambda t, q, k: O.knn_kdtree(t, q, k, workers=-1)
    xt = O.acos_threshold(math.pi * 5 / 12)
    d = np.float32(2) * O.average_edge_length(noisy, knn(noisy, noisy, 6))
    pos = noisy
    pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, knn=knn)       # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, knn=knn)
    dt = time.perf_counter() - t0
    return n_sample * iters / dt, cores, dt / iters


def run_reference(args):
    """--impl reference: the CPU implementation (oracle port; the reference itself is Python and cannot travel)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.ref_points
    import numpy as np
    import torch
    import ngpd_oracle as O
    from ngpd_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clean, normal = workloads.creased_surface(n_sample, 4321, "cpu")
    noisy = workloads.add_noise(clean, 0.3 * workloads.expected_spacing(n_sample)).numpy()
    nrm = normal.numpy()
    knn = lambda t, q, k: O.knn_kdtree(t, q, k, workers=-1)
    xt = O.acos_threshold(math.pi * 5 / 12)
    d = np.float32(2) * O.average_edge_length(noisy, knn(noisy, noisy, 6))
    pos = noisy
    for _ in range(args.warmup):
        pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, knn=knn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pos, nrm, _, _ = O.denoise_iteration(noisy, pos, nrm, K_F, K_U, xt, ALPHAS, d, knn=knn)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = (f"{n_sample}-point cloud from the same generator (the full workload has {args.points} points; throughput per point is "
              f"size-independent to first order for a KD-tree pipeline), {args.steps} timed iterations of the oracle port "
              f"(NumPy + SciPy cKDTree workers=-1 + LAPACK via torch, all {cores} host threads)")
    line = {"impl": "reference", "metric": "denoise point-iterations/sec (kNN+NVT+update)", "value": value, "unit": "point-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.points),
            "cpu_baseline": {"value": value, "unit": "point-iterations/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "point-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n):
    name = ("BASELINE.json configs[3]: synthetic 100M-point noisy CAD-like surface (cube faces 60 % + torus 40 %, isotropic Gaussian noise "
            f"sigma = 0.3 x mean 6-NN distance, shuffled), k={K_F}/{K_U}; one Processor.denoise iteration body per step; the whole cloud "
            "fits one B200 (37 GB), more GPUs split the same cloud into Morton slabs")
    if K_F == 32:
        name = ("BASELINE.json configs[2] stand-in (xyzrgb_dragon is a missing blob, SURVEY 8d): the synthetic creased surface at "
                f"{n} points, Gaussian noise, k={K_F}/{K_U}, one Processor.denoise iteration body per step")
    return {"workload": name,
            "points": n, "k_feature": K_F, "k_update": K_U, "strategy": "flat/edge/feature", "alpha": list(ALPHAS),
            "l2": "inputs larger than L2 (positions+normals+neighbour table >> 126 MB); no flush needed",
            "partition": "single GPU" if args.gpus == 1 else f"{args.gpus} Morton slabs + halo exchange"}


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

    if world == 1:
        noisy, nrm = make_input(n, dev)
        sess = _lib.Session(noisy, K_F)
        sess.set_state(noisy, nrm)
        s, c = sess.mean_edge_length_parts(6)
        params = _lib.make_params(K_F, K_U, None, 0.3, 3.0, 0.2, (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE), ALPHAS, 2.0 * s / c)
        step = lambda: sess.step(params)
        profile_src = sess
        n_local = n
    else:
        from ngpd_b200 import partition
        noisy, nrm = make_input(n, dev)                      # replicated deterministic input; each rank keeps its slab
        slab = partition.SlabSession(noisy, nrm, K_F, K_U, ALPHAS)
        del noisy, nrm
        step = slab.step
        profile_src = slab.session
        n_local = slab.n_owned

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    profile_src.set_profiling(True)
    profile_src.get_profile()
    launches_before = profile_src.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = profile_src.get_profile()
    profile_src.set_profiling(False)
    # single GPU: the fused step restarts its counter, so the last step's count x steps; slabs: the phase calls accumulate
    launches = profile_src.launch_count() * args.steps if world == 1 else profile_src.launch_count() - launches_before
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    value = n * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer entry point: pinned host -> device, one iteration, device -> host
    e2e = None
    # (pinned buffers are allocated on the GPU's own NUMA node: partition.near_gpu)
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
    else:
        # every rank streams its own slab through pinned host buffers; halo rows come from their owners over NCCL
        k = slab.n_owned
        _, p, q, _ = slab.owned_state()
        pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
        with partition.near_gpu(dev.index) as cpus:
            pos_h, nrm_h, pos_o, nrm_o, lab_o = pin(k, 3), pin(k, 3), pin(k, 3), pin(k, 3), pin(k, dtype=torch.uint8)
        print(f"[bench] rank {rank}: pinned staging allocated on the CPUs next to cuda:{dev.index}: {len(cpus)} of {os.cpu_count()}", file=sys.stderr)
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
               "steps": e_steps, "call": "SlabSession.step_host (each rank: pinned host buffers of its slab in and out, halo exchange over NCCL)"}

    # ---- the metric's second half: kNN queries/s through the public ngpd_knn (no temporal coherence: every query searched)
    knn_line = None
    if world == 1 and not args.no_knn:
        del sess
        profile_src = None
        torch.cuda.empty_cache()
        grid = _lib.Grid(noisy, K_F)
        flags = _lib.KNN_QUERY_IS_TREE
        for _ in range(2):
            tab = grid.knn(noisy, K_F, flags)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(3):
            tab = grid.knn(noisy, K_F, flags)
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / 3
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2 = grid.nn_sqdist(noisy, False, flags)
        c0.record()
        for _ in range(3):
            d2 = grid.nn_sqdist(noisy, False, flags)
        c1.record()
        torch.cuda.synchronize()
        cms = c0.elapsed_time(c1) / 3
        knn_line = {"metric": "kNN queries/sec", "k": K_F, "value": n / (kms * 1e-3), "ms": kms, "algorithmic_bytes_per_query": 24 + 4 * K_F,
                    "frac_of_hbm_peak": (24 + 4 * K_F) * n / (kms * 1e-3) / 1e9 / peak,
                    "nearest_neighbour_queries_per_s": n / (cms * 1e-3), "nearest_ms": cms,
                    "call": "ngpd_knn / ngpd_nn_sqdist (public ABI, queries = the tree's own noisy points, index output materialised)"}
        del tab, d2, grid

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    kernels = {}
    for name, (tms, cnt) in prof.items():
        if cnt == 0:
            continue
        per_step = tms / args.steps
        ab = bytes_per[name] * n_local
        kernels[name] = {"ms_per_step": per_step, "launches_per_step": cnt / args.steps, "algorithmic_bytes_per_point": bytes_per[name],
                         "achieved_gbs": ab / (per_step * 1e-3) / 1e9, "frac": ab / (per_step * 1e-3) / 1e9 / peak,
                         "share_of_step": per_step / (ms / args.steps)}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    traffic, traffic_src = ncu_traffic(n_local)
    for name in kernels:
        kernels[name]["ncu_dram_bytes_per_step"] = traffic.get(name)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "kNN is instruction/latency-bound (register top-k, fp64 distances); reported against HBM as SURVEY 8(d) asks",
                "iteration_achieved_gbs": bytes_per["iteration"] * n / (ms / args.steps * 1e-3) / 1e9,
                "iteration_frac": bytes_per["iteration"] * n / (ms / args.steps * 1e-3) / 1e9 / peak}
    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, per_it = cpu_iteration_rate(args.cpu_points, 5)          # ~11 s of CPU work at 200 k points
        cpu = {"value": v, "unit": "point-iterations/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_points}-point cloud from the same generator, 1 warm-up + 5 timed iterations of the oracle port "
                         f"(NumPy + SciPy KD-tree workers=-1 + LAPACK), {per_it:.2f} s per iteration"}
    line = {"metric": "denoise point-iterations/sec (kNN+NVT+update)", "value": value, "unit": "point-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n), "clocks": clocks,
            "e2e": e2e, "knn": knn_line, "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k-feature", type=int, default=K_F, help="k of the feature pass (configs[2]: 32)")
    ap.add_argument("--k-update", type=int, default=K_U)
    args = ap.parse_args()
    K_F, K_U = args.k_feature, args.k_update
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as ge
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            ge.build()
        run_ours(args)


if __name__ == "__main__":
    main()
