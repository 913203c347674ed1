"""Golden vectors for the neighbourhood size of BASELINE configs[2] (k = 32) from the UNMODIFIED reference (run in the build
container only; same stub-import recipe as make_golden.py):

    python tests/golden/make_golden_k32.py            # writes tests/golden/fandisk_k32.npz

Recorded, on models/fandisk_gaus_n6_noisy.obj with the oriented PCA normals that fandisk_denoise.npz holds (`pos0`, `n_flip`,
themselves reference outputs): Processor.getMyFeatureDecomposition(N=32) (Processor.py:110-117) -- the 32-NN table, both voting
tensors with their eigenpairs, the smoothed normals, labels -- and the three class steps with the 8-NN selection as in
Processor.denoise (:124-138).  Nothing here is computed by this repository's own code."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NGPD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(1, REF)

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
npf = lambda t: t.detach().cpu().numpy()


def main():
    K = 32
    base = np.load(os.path.join(HERE, "fandisk_denoise.npz"))
    p = Processor(Pointcloud(torch.from_numpy(base["pos0"].copy())))
    g = p.graph
    g.n = torch.from_numpy(base["n_flip"].copy())
    out = {"k": np.int64(K)}
    l = TorchUtils.averageEdgeLength(g.pos, p.selector.getKNNSelection(6).getEdgeIndex())
    d = 2 * l
    out["l"] = np.float32(l.item())
    dec, f_n = p.getMyFeatureDecomposition(K)
    sel = p.selector.getKNNSelection(K)
    out["knn32"] = npf(sel.j).reshape(-1, K).astype(np.int32)
    out["T1"], out["T2"] = npf(_tensors[0]), npf(_tensors[1])
    nvt1 = p.decompositionor.getBetterFilteredNVT(sel, g.n, torch.pi * 5 / 12)
    out["eigval1"], out["eigvec1"] = npf(nvt1.eigval), npf(nvt1.eigvec)
    out["f_n"] = npf(f_n).copy()
    out["eigval2"], out["eigvec2"] = npf(dec.eigval), npf(dec.eigvec)
    classes = dec.getClasses()
    out["classes"] = npf(classes).astype(np.uint8)
    sel8 = p.selector.getKNNSelection(8)
    alphas = [1, 0.2, 1]
    for key in range(3):
        idx = (classes == key).nonzero().flatten()
        if idx.size(0) == 0:
            continue
        if key == 0:
            new = p.denoiser.flat_step(sel8.filter(idx), f_n, d, alphas[key])
        elif key == 1:
            new = p.denoiser.edge_step(sel8.filter(idx), f_n, dec.eigvec[..., 0], d, alphas[key])
        else:
            new = p.denoiser.feature_step(sel8.filter(idx), f_n, d, alphas[key])
        g.pos[idx] = new
    out["pos_after"] = npf(g.pos).copy()
    np.savez_compressed(os.path.join(HERE, "fandisk_k32.npz"), **out)
    print("fandisk_k32.npz", "N", len(out["knn32"]), "labels", np.bincount(out["classes"]))


if __name__ == "__main__":
    main()
