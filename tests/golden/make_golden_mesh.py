"""Golden vectors of the mesh vertex update from the UNMODIFIED reference class
PatchGeneration.Modules.Mesh (run in the build container only):

    python tests/golden/make_golden_mesh.py            # writes tests/golden/mesh_update.npz

models/fandisk.obj with Gaussian vertex noise, target normals = the clean mesh's face normals, Mesh.updateVertices(n, k)
for k = 1 and k = 5 (Mesh.py:377-418).  igl / polyscope / meshplot are not installed here: the adjacency that
igl.vertex_triangle_adjacency would return is passed to the constructor (its documented layout), the other two are never
called on this path.  Nothing below is computed by this repository's own code except that adjacency and the OBJ parsing."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NGPD_REFERENCE", "/root/reference")
for name in ("polyscope", "meshplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstubs"))
sys.path.insert(1, REF)

from PatchGeneration.Modules.Mesh import Mesh  # noqa: E402


def read_obj(path):
    v, f = [], []
    for line in open(path):
        if line.startswith("v "):
            v.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            f.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int64)


def vta(f, n):
    face = np.repeat(np.arange(len(f)), 3)
    vert = f.reshape(-1)
    order = np.lexsort((face, vert))
    ni = np.zeros(n + 1, dtype=np.int64)
    ni[1:] = np.cumsum(np.bincount(vert, minlength=n))
    return face[order], ni


def main():
    v, f = read_obj(os.path.join(REF, "models", "fandisk.obj"))
    adj = vta(f, len(v))
    clean = Mesh(v.copy(), f, f2f=np.zeros((len(f), 3), dtype=np.int64), vta=adj)
    n = clean.getFaceNormals()
    rng = np.random.default_rng(3)
    fv = v[f]
    edge = np.linalg.norm(fv[:, 1] - fv[:, 0], axis=1).mean()
    noisy = v + rng.normal(0, 0.2 * edge, v.shape)
    out = {"v_noisy": noisy, "f": f.astype(np.int32), "face_normals": n, "vta_faces": adj[0].astype(np.int32), "vta_offsets": adj[1].astype(np.int32)}
    for k in (1, 5):
        m = Mesh(noisy.copy(), f, f2f=np.zeros((len(f), 3), dtype=np.int64), vta=adj)
        m.updateVertices(n, k)
        out[f"v_after_{k}"] = m.getVertices().copy()
    np.savez_compressed(os.path.join(HERE, "mesh_update.npz"), **out)
    d0 = np.linalg.norm(noisy - v, axis=1).mean(); d5 = np.linalg.norm(out["v_after_5"] - v, axis=1).mean()
    print("mesh_update.npz V", len(v), "F", len(f), "mean vertex error", d0, "->", d5)


if __name__ == "__main__":
    main()
