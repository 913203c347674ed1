"""Yardstick for the free-running end-to-end comparison: the oracle (same LAPACK as the reference) run on the recorded
fandisk input, and again with its input normals moved by 1 ulp.  Prints how far the reference's algorithm differs from
ITSELF after Processor.denoise()'s two iterations (positions, normals, labels, Chamfer mean).  Run in the build container."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ngpd_oracle as O
from conftest import angle_between
f = dict(np.load(os.path.join(ROOT, "tests", "golden", "fandisk_denoise.npz")))
pos0, n0, gt = f["pos0"], f["n_flip"], f["gt"]
pa, na, la = O.denoise(pos0, pos0, n0)
for trial, seed in enumerate((0, 1, 2)):
    rng = np.random.default_rng(seed)
    n1 = np.nextafter(n0, np.where(rng.random(n0.shape) < 0.5, -2, 2).astype(np.float32))
    pb, nb, lb = O.denoise(pos0, pos0, n1)
    scale = np.abs(pa).max()
    err = np.abs(pa - pb).max(axis=1) / scale
    cda = O.chamfer_distance(gt, pa).mean(dtype=np.float64); cdb = O.chamfer_distance(gt, pb).mean(dtype=np.float64)
    print(f"trial {trial}: positions >1e-5: {(err > 1e-5).mean():.4%} (max {err.max():.2e}); normals >1e-4 rad: "
          f"{(angle_between(na, nb) > 1e-4).mean():.4%}; labels differ: {(la != lb).sum()}; CD {cda:.6e} vs {cdb:.6e} rel {abs(cda - cdb) / cda:.2e}")
ref = f["cd_final"].mean(dtype=np.float64)
print("oracle vs recorded reference CD rel:", abs(O.chamfer_distance(gt, pa).mean(dtype=np.float64) - ref) / ref)

# ---- the until-minimum-error loop (BASELINE config 2, recorded bunny): strategy flat/feature/feature, k = 8, alpha (1,.2,1), d = 2l
u = dict(np.load(os.path.join(ROOT, "tests", "golden", "until_min.npz")))
its = int(u["iterations"]) + 1          # the loop body ran once more than the returned count
xt = O.acos_threshold(5 * np.pi / 12)


def run(n_in):
    pos, nrm = u["pos0"], n_in
    for _ in range(its):
        pos, nrm, _, _ = O.denoise_iteration(u["pos0"], pos, nrm, 16, 8, xt, (1.0, 0.2, 1.0), np.float32(2) * u["l"],
                                             strategy=("flat", "feature", "feature"))
    return pos
base = run(u["n_flip"])
scale = np.abs(u["pos0"]).max()
print("until-min: oracle vs recorded reference positions >1e-5:", (np.abs(base - u["pos_returned"]).max(axis=1) / scale > 1e-5).mean())
for seed in (0, 1, 2):
    rng = np.random.default_rng(seed)
    n1 = np.nextafter(u["n_flip"], np.where(rng.random(u["n_flip"].shape) < 0.5, -2, 2).astype(np.float32))
    err = np.abs(run(n1) - base).max(axis=1) / scale
    print(f"until-min trial {seed}: after {its} iterations positions >1e-5: {(err > 1e-5).mean():.4%} (max {err.max():.2e})")
