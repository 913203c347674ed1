"""Per-step kernel-group times and k-NN tier counts on two inputs: round 1's generator and round 2's chunked one."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import ngpd_b200
from ngpd_b200 import _lib, workloads as W
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)


def run(name, noisy, nrm, steps=8):
    sess = _lib.Session(noisy, 16)
    sess.reserve(16)
    sess.set_state(noisy, nrm)
    s, c = sess.mean_edge_length_parts(6)
    params = _lib.make_params(dmax=2.0 * s / c)
    sess.set_profiling(True)
    for it in range(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sess.step(params)
        e1.record()
        torch.cuda.synchronize()
        prof = sess.get_profile()
        st = sess.knn_stats()
        d = sess.checksum()
        print(f"{name} it {it}: {e0.elapsed_time(e1):.3f} ms  " + " ".join(f"{k}={v[0]:.3f}" for k, v in prof.items() if v[1]) +
              f"  tiers(t0->search, t1->t2, t2->exact)={st}  labels={d[6:9]} hash={d[0]:016x}/{d[1]:016x}", flush=True)


args = argparse.Namespace(surface="creased", strategy="flat/edge/feature", clamp=False)
noisy, analytic, _ = bench.make_shard(args, n, dev, 0, 1)
nrm = bench.single_gpu_normals(noisy, analytic)
run("chunked", noisy, nrm)
del noisy, nrm, analytic
sys.exit(0)
sess = _lib.Session(clean, 16)
s, c = sess.mean_edge_length_parts(6)
del sess
noisy = W.add_noise(clean, 0.3 * s / c * 6.0 / 5.0)
nrm = bench.single_gpu_normals(noisy, normal)
pass
