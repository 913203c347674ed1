"""BASELINE.json configs[0] and configs[1] through the mirror of the reference API, on the clouds recorded in tests/golden/
(the GPU box has no /root/reference): wall time of Processor.denoise() and of denoiseUntilMinimumError incl. the Chamfer
evaluation each iteration, next to the reference's own time for the same call (measured at survey
time in the build container).  Markdown on stdout."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ngpd_b200 as ng

G = os.path.join(ROOT, "tests", "golden")
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def wall(fn, reps):
    """median wall time of fn() (host API calls included: a Processor allocates ~40 device buffers), after two warm-up calls"""
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], out


print("| config | cloud | points | call | ours ms | iterations | point-iterations/s | CD before -> after | reference (CPU, survey) |")
print("|---|---|---|---|---|---|---|---|---|")
f = dict(np.load(os.path.join(G, "fandisk_denoise.npz")))
n = len(f["pos0"])


def config1():
    p = ng.Processor(ng.Pointcloud(cu(f["pos0"]).clone()))
    p.graph.n = cu(f["n_flip"]).clone()
    p.denoise()
    return p.graph.pos


t, pos = wall(config1, 21)
cd0 = float(ng.TorchUtils.ChamferDistance(cu(f["gt"]), cu(f["pos0"])).mean()); cd1 = float(ng.TorchUtils.ChamferDistance(cu(f["gt"]), pos).mean())
print(f"| 0 | models/fandisk_gaus_n6_noisy.obj | {n} | Processor(pc) + denoise() (2 iterations, k=16/8) | {t * 1e3:.2f} | 2 | {2 * n / t:.3g} | {cd0:.4f} -> {cd1:.4f} | 1.77 s cold, CD 0.3853 -> 0.2264 |")

P0 = ng.Processor(ng.Pointcloud(cu(f["pos0"]).clone()))


def config1_resident():
    P0.graph.pos.copy_(cu(f["pos0"])); P0.graph.n = cu(f["n_flip"]).clone()
    P0.denoise()
    return P0.graph.pos


t, _ = wall(config1_resident, 21)
print(f"| 0 | (same, Processor built once) | {n} | denoise() (2 iterations, k=16/8) incl. upload of the cloud | {t * 1e3:.2f} | 2 | {2 * n / t:.3g} | | |")

u = dict(np.load(os.path.join(G, "until_min.npz")))
n = len(u["pos0"])
iters = {}


def config2():
    p = ng.Processor(ng.Pointcloud(cu(u["pos0"]).clone()))
    p.graph.n = cu(u["n_flip"]).clone()
    strategy = {0: p.denoiser.flat_step, 1: p.denoiser.feature_step, 2: p.denoiser.feature_step}
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        best, err, it = p.denoiseUntilMinimumError(cu(u["gt"]), strategy, k=8, alpha=[1, 0.2, 1], d=2 * float(u["l"]),
                                                    error_funcs=[ng.TorchUtils.ChamferDistance])
    iters["n"] = it
    return best, err


t, (best, err) = wall(config2, 11)
it = iters["n"] + 1
cd0 = float(ng.TorchUtils.ChamferDistance(cu(u["gt"]), cu(u["pos0"])).mean())
print(f"| 1 | Generated_Noise/{str(u['cloud'])}.obj | {n} | Processor(pc) + denoiseUntilMinimumError (flat/feature/feature, k=8, CD each iteration) | {t * 1e3:.2f} | {it} | "
      f"{it * n / t:.3g} | {cd0:.4e} -> {float(err[0].mean()):.4e} | {it} iterations, 35.3 k point-iterations/s excl. Chamfer, CD 1.7154e-7 -> 1.5591e-7 |")

# ---- BASELINE configs[2]'s preparation step on the device: a mesh upsampled to 10 M points (Pointcloud.sampleObj, Object.py:135-156),
# PCA normals over the 12-NN graph, spanning-tree orientation, Gaussian noise along the normals (Noise.generateNoise, Noise.py:33-59).
# xyzrgb_dragon is a missing blob: a torus mesh of 200 x 100 quads stands in.
import math, tempfile
nu, nv = 200, 100
uu, vv = np.meshgrid(np.arange(nu) * 2 * math.pi / nu, np.arange(nv) * 2 * math.pi / nv, indexing="ij")
verts = np.stack([(0.6 + 0.25 * np.cos(vv)) * np.cos(uu), (0.6 + 0.25 * np.cos(vv)) * np.sin(uu), 0.25 * np.sin(vv)], -1).reshape(-1, 3)
vid = lambda i, j: (i % nu) * nv + (j % nv) + 1
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "torus.obj")
    with open(path, "w") as fh:
        for v in verts:
            fh.write(f"v {v[0]:.7f} {v[1]:.7f} {v[2]:.7f}\n")
        for i in range(nu):
            for j in range(nv):
                fh.write(f"f {vid(i, j)} {vid(i + 1, j)} {vid(i + 1, j + 1)}\nf {vid(i, j)} {vid(i + 1, j + 1)} {vid(i, j + 1)}\n")
    m = 10_000_000
    torch.manual_seed(0); torch.cuda.manual_seed(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pc = ng.Pointcloud.sampleObj(path, m, device="cuda")
    torch.cuda.synchronize(); t_sample = time.perf_counter() - t0
p = ng.Processor(pc)
torch.cuda.synchronize(); t0 = time.perf_counter()
p.graph.edge_index = p.graphBuilder.getKNNEdgeIndex(12)
torch.cuda.synchronize(); t_graph = time.perf_counter() - t0; t0 = time.perf_counter()
p.graphBuilder.setAndFlipNormals(flip=True)
torch.cuda.synchronize(); t_normals = time.perf_counter() - t0; t0 = time.perf_counter()
l = ng.TorchUtils.averageEdgeLength(p.graph.pos, p.graph.edge_index)
p.noise.generateNoise(0.3, float(l), keepNormals=True)
torch.cuda.synchronize(); t_noise = time.perf_counter() - t0
agree = float(((p.graph.n * pc.n).sum(1).abs() > 0.9).float().mean())
flipped = float(((p.graph.n * pc.n).sum(1) > 0).float().mean())
print(f"\nPreparation of a 10 M-point cloud on the device (configs[2]'s upsampling; torus mesh of {2 * nu * nv} triangles): sampleObj {t_sample * 1e3:.0f} ms "
      f"(incl. parsing the OBJ), 12-NN graph {t_graph * 1e3:.0f} ms, PCA normals + spanning-tree orientation {t_normals * 1e3:.0f} ms "
      f"({p.graphBuilder.orientation_info}), noise {t_noise * 1e3:.0f} ms; PCA normals within 26 deg of the face normals on {agree:.4%} of the points, "
      f"{max(flipped, 1 - flipped):.4%} consistently oriented")
