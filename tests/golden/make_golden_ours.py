"""Golden vectors of the thesis' final strategy ("Ours") and of "CTD-QEM" from the UNMODIFIED reference (run in the build
container only; same stub-import recipe as make_golden.py):

    python tests/golden/make_golden_ours.py            # writes tests/golden/ours_fandisk.npz

Recorded, for the loops of PostProcessing.ipynb#c9 on models/fandisk_gaus_n6_noisy.obj:
  "Ours"    (j == 3): 2 iterations; getMyFeatureDecomposition, getClasses, flat_step for class 0 and feature_step for classes
            1 and 2 with d * 20000 -- all read graph.pos while the results go to temp_pos (every class moves from the same
            snapshot) --, then the displacement clamp  mask = |temp_pos - original_pos| < d;
  "CTD-QEM" (j == 2): 5 iterations of feature_step on every point.
Nothing here is computed by this repository's own code."""
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

from Pointcloud.Modules.Object import Pointcloud  # noqa: E402
from Pointcloud.Modules.Processor import Processor  # noqa: E402
from Pointcloud.Modules.Utils import TorchUtils  # noqa: E402


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
    original_n = g.n.clone()
    # ---- "Ours"
    alphas = [1, 0.2, 1]
    for it in range(2):
        tag = f"ours{it}_"
        out[tag + "pos_in"] = npf(g.pos).copy(); out[tag + "n_in"] = npf(g.n).copy()
        decomposition, f_n = p.getMyFeatureDecomposition()
        classes = decomposition.getClasses()
        selection = p.selector.getKNNSelection(8)
        temp_pos = g.pos.clone()
        for key in range(3):
            indices = (classes == key).nonzero().flatten()
            if indices.size(0) == 0:
                continue
            elif key == 0:
                new_pos = p.denoiser.flat_step(selection.filter(indices), f_n, d * 20000, alphas[key])
            else:
                new_pos = p.denoiser.feature_step(selection.filter(indices), f_n, d * 20000, alphas[key])
            temp_pos[indices] = new_pos
        mask = (temp_pos - original_pos).norm(dim=1) < d
        g.pos[mask] = temp_pos[mask]
        g.n = f_n
        out[tag + "f_n"] = npf(f_n).copy(); out[tag + "classes"] = npf(classes).astype(np.uint8)
        out[tag + "temp_pos"] = npf(temp_pos).copy(); out[tag + "mask"] = npf(mask); out[tag + "pos_out"] = npf(g.pos).copy()
    g.pos = original_pos.clone()
    g.n = original_n.clone()
    # ---- "CTD-QEM"
    for it in range(5):
        tag = f"qem{it}_"
        out[tag + "pos_in"] = npf(g.pos).copy(); out[tag + "n_in"] = npf(g.n).copy()
        _, f_n = p.getMyFeatureDecomposition()
        selection = p.selector.getKNNSelection(8)
        g.pos = p.denoiser.feature_step(selection, f_n, d, 1)
        g.n = f_n
        out[tag + "pos_out"] = npf(g.pos).copy(); out[tag + "f_n"] = npf(f_n).copy()
    np.savez_compressed(os.path.join(HERE, "ours_fandisk.npz"), **out)
    print("ours_fandisk.npz N", len(out["pos0"]), "classes", [np.bincount(out[f"ours{i}_classes"], minlength=3).tolist() for i in range(2)],
          "clamped", [int((~out[f"ours{i}_mask"]).sum()) for i in range(2)])


if __name__ == "__main__":
    th = threading.Thread(target=main)
    th.start()
    th.join()
