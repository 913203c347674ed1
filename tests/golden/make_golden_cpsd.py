"""Golden vectors of the Yadav-2018 baseline path ("CPSD") from the UNMODIFIED reference
(run in the build container only; same stub-import recipe as make_golden.py):

    python tests/golden/make_golden_cpsd.py            # writes tests/golden/cpsd_fandisk.npz

Recorded, for the loop of PostProcessing.ipynb#c9 (method "CPSD", j == 1) on models/fandisk_gaus_n6_noisy.obj:
radius-ball selection (Selector.py:214-233), getNormalFilteredNVT / getNormalFilteredPVT (Decompositionor.py:260-276,
172-211) with their eigenpairs, smoothed normals, getVUFeatures labels, the three class steps on CSR rows and the
accepted positions, for 3 iterations -- plus one call of each step kind on ball rows with a synthetic class split so
that every kind sees variable-length rows.  Nothing here is computed by this repository's own code."""
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NGPD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(1, REF)
sys.setrecursionlimit(10_000_000)
threading.stack_size(1024 * 1024 * 1024)

import torch  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)

from Pointcloud.Modules import Decompositionor as ref_dec  # noqa: E402
from Pointcloud.Modules.Object import Pointcloud  # noqa: E402
from Pointcloud.Modules.Processor import Processor  # noqa: E402
from Pointcloud.Modules.Utils import TorchUtils  # noqa: E402

_tensors = []
_orig_eigh = ref_dec.torch_linalg_eigh


def _capturing_eigh(T):
    _tensors.append(T.clone())
    return _orig_eigh(T)


ref_dec.torch_linalg_eigh = _capturing_eigh


def npf(t):
    return t.detach().cpu().numpy()


def main():
    base = dict(np.load(os.path.join(HERE, "fandisk_denoise.npz")))
    p = Processor(Pointcloud.loadObj(os.path.join(REF, "models", "fandisk_gaus_n6_noisy.obj")))
    g = p.graph
    g.n = torch.from_numpy(base["n_flip"].copy())
    out = {"pos0": npf(g.pos).copy(), "n0": npf(g.n).copy()}
    l = TorchUtils.averageEdgeLength(g.pos, p.selector.getKNNSelection(6).getEdgeIndex())
    d = 2 * l
    out["l"] = np.float32(l.item()); out["d"] = np.float32(d.item())
    original_pos = g.pos.clone()
    alphas = [0.1, 1, 1]
    for it in range(3):
        tag = f"it{it}_"
        out[tag + "pos_in"] = npf(g.pos).copy(); out[tag + "n_in"] = npf(g.n).copy()
        _tensors.clear()
        sel = p.selector.getPointsInRangeSelection(d)
        out[tag + "ball_j"] = npf(sel.j).astype(np.int32); out[tag + "ball_slices"] = npf(sel.slices).astype(np.int64)
        dec, f_n = p.getMartinFeatureDecomposition(r=d)
        out[tag + "T_nvt"] = npf(_tensors[0]); out[tag + "T_pvt"] = npf(_tensors[1])
        nvt = p.decompositionor.getNormalFilteredNVT(sel, g.n, 0.9)
        out[tag + "nvt_eigval"] = npf(nvt.eigval); out[tag + "nvt_eigvec"] = npf(nvt.eigvec)
        out[tag + "f_n"] = npf(f_n).copy()
        out[tag + "pvt_eigval"] = npf(dec.eigval); out[tag + "pvt_eigvec"] = npf(dec.eigvec)
        classes = dec.getVUFeatures(tau=0.3)
        out[tag + "classes"] = npf(classes).astype(np.uint8)
        # the same decomposition at a scale-aware threshold, so that all three labels occur (tau = 4 l^2)
        tau2 = float(4 * l * l)
        out[tag + "classes_tau2"] = npf(dec.getVUFeatures(tau=tau2)).astype(np.uint8); out["tau2"] = np.float32(tau2)
        sel8 = p.selector.getKNNSelection(k=8)
        temp_pos = g.pos.clone()
        for key in range(3):
            idx = (classes == key).nonzero().flatten()
            if idx.size(0) == 0:
                continue
            if key == 0:
                new = p.denoiser.flat_step(sel8.filter(idx), f_n, d * 20000, alphas[key])
            elif key == 1:
                new = p.denoiser.edge_step(sel8.filter(idx), f_n, dec.eigvec[..., 0], d * 20000, alphas[key])
            else:
                new = p.denoiser.corner_step(sel8.filter(idx), f_n, d * 20000, alphas[key])
            temp_pos[idx] = new
        out[tag + "temp_pos"] = npf(temp_pos).copy()
        mask = (temp_pos - original_pos).norm(dim=1) < d
        g.pos[mask] = temp_pos[mask]
        g.n = f_n
        out[tag + "pos_out"] = npf(g.pos).copy()
        if it == 0:
            # every step kind on variable-length ball rows, synthetic class split i % 3 (teacher-forced inputs)
            n = g.pos.size(0)
            pos_in = torch.from_numpy(out["it0_pos_in"].copy())
            keep = g.pos.clone()
            g.pos = pos_in
            for key, name in enumerate(("flat", "edge", "corner", "feature")):
                idx = (torch.arange(n) % 4 == key).nonzero().flatten()
                rows = sel.filter(idx)
                if name == "flat":
                    new = p.denoiser.flat_step(rows, f_n, d, 0.5)
                elif name == "edge":
                    new = p.denoiser.edge_step(rows, f_n, dec.eigvec[..., 0], d, 0.5)
                elif name == "corner":
                    new = p.denoiser.corner_step(rows, f_n, d, 0.5)
                else:
                    new = p.denoiser.feature_step(rows, f_n, d, 0.5)
                out["ballrows_" + name] = npf(new).copy()
            g.pos = keep
    cnt = np.diff(out["it0_ball_slices"])
    np.savez_compressed(os.path.join(HERE, "cpsd_fandisk.npz"), **out)
    print("cpsd_fandisk.npz N", len(out["pos0"]), "ball sizes min/mean/max", cnt.min(), cnt.mean(), cnt.max(),
          "labels", [np.bincount(out[f"it{i}_classes"], minlength=3).tolist() for i in range(3)],
          "labels tau2", np.bincount(out["it0_classes_tau2"], minlength=3).tolist())


if __name__ == "__main__":
    th = threading.Thread(target=main)
    th.start()
    th.join()
