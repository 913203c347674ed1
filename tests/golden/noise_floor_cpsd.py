"""Yardstick for the free-running comparison of the Yadav-2018 baseline loop ("CPSD", PostProcessing.ipynb#c9): the oracle (same
LAPACK as the reference) run for three iterations on the recorded fandisk input, and again with its input normals moved by
1 ulp.  Prints how far the reference's algorithm differs from ITSELF per iteration (labels, normals, positions): the floor
tests/test_gpu_parity.py::test_cpsd_loop_vs_reference is read against.  Run in the build container."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ngpd_oracle as O
from conftest import angle_between
c = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpsd_fandisk.npz")))
pos0, n0, d = c["pos0"], c["n0"], c["d"]


def run(n_in):
    pos, nrm, out = pos0.copy(), n_in.copy(), []
    for it in range(3):
        pos, nrm, lab = O.cpsd_iteration(pos0, pos, nrm, pos0, d)[:3]
        out.append((pos.copy(), nrm.copy(), lab.copy()))
    return out


base = run(n0)
scale = np.abs(pos0).max()
for it in range(3):
    t = f"it{it}_"
    e = np.abs(base[it][0] - c[t + "pos_out"]).max(axis=1) / scale
    print(f"oracle vs recorded reference, iteration {it}: labels agree {(base[it][2] == c[t + 'classes']).mean():.4%}, normals > 1e-4 rad "
          f"{(angle_between(base[it][1], c[t + 'f_n']) > 1e-4).mean():.4%}, positions > 1e-5 {(e > 1e-5).mean():.4%}")
for seed in (0, 1, 2):
    rng = np.random.default_rng(seed)
    n1 = np.nextafter(n0, np.where(rng.random(n0.shape) < 0.5, -2, 2).astype(np.float32))
    other = run(n1)
    for it in range(3):
        e = np.abs(base[it][0] - other[it][0]).max(axis=1) / scale
        print(f"1-ulp trial {seed}, iteration {it}: labels agree {(base[it][2] == other[it][2]).mean():.4%}, normals > 1e-4 rad "
              f"{(angle_between(base[it][1], other[it][1]) > 1e-4).mean():.4%}, positions > 1e-5 {(e > 1e-5).mean():.4%} (max {e.max():.2e})")
