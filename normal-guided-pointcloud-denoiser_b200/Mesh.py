"""The part of the reference's PatchGeneration/Modules/Mesh.py that the Vertex_updating notebook exercises
(SURVEY.md 8f rank 4): a triangle mesh with its vertex-triangle adjacency and `updateVertices(n, k)` -- move the
vertices so that the face normals align with given target normals (Mesh.py:377-418).  numpy arrays in and out, fp64,
like the reference; the sweeps run on the GPU through ngpd_mesh_vertex_update."""
from __future__ import annotations

import errno
import os

import numpy as np
import torch

from . import _lib
from .Object import read_obj


def vertex_triangle_adjacency(f: np.ndarray, n: int):
    """igl.vertex_triangle_adjacency(F, n): (VF, NI) -- the faces incident to each vertex, concatenated vertex by vertex
    in ascending face order, and the n+1 offsets into that list."""
    f = np.asarray(f, dtype=np.int64)
    face = np.repeat(np.arange(len(f), dtype=np.int64), 3)
    vert = f.reshape(-1)
    order = np.lexsort((face, vert))
    ni = np.zeros(n + 1, dtype=np.int64)
    ni[1:] = np.cumsum(np.bincount(vert, minlength=n))
    return face[order], ni


class Mesh:
    def __init__(self, v, f, noise_factor=0, f2f=None, vta=None, gt=None):
        self.v = np.asarray(v, dtype=np.float64)
        self.f = np.asarray(f, dtype=np.int64)
        self.noise_factor = noise_factor
        self.f2f = f2f
        self.vta = vta if vta is not None else vertex_triangle_adjacency(self.f, len(self.v))
        self.gt = gt

    @classmethod
    def readFile(cls, obj_file):
        if not type(obj_file) == str:
            raise ValueError("obj_file (first argument) must be a string representing the path towards the object file.")
        if not obj_file.endswith(".obj"):
            raise ValueError("obj_file (first argument) must be a path towards and object file ending with '.obj'")
        if not os.path.exists(obj_file):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), obj_file)
        v, _, f, _ = read_obj(obj_file)
        return Mesh(v, f)

    def getVertices(self):
        return self.v

    def getFaceNormals(self):
        fv = self.getVertices()[self.f]
        crosses = np.cross(fv[:, 1, :] - fv[:, 0, :], fv[:, 2, :] - fv[:, 1, :])
        return crosses / np.linalg.norm(crosses, axis=1)[:, None]

    def getVertexTriangleAdjacency(self):
        return self.vta

    def updateVertices(self, n, k=15):
        """k Jacobi sweeps of  v_i += 1/(3 deg_i) sum_{f at i} sum_{c in f} n_f (n_f . (v_c - v_i))  (Mesh.py:377-418);
        updates self.v in place like the reference."""
        _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        v = torch.from_numpy(np.ascontiguousarray(self.getVertices(), dtype=np.float64)).to(dev)
        f = torch.from_numpy(np.ascontiguousarray(self.f, dtype=np.int32)).to(dev)
        fn = torch.from_numpy(np.ascontiguousarray(n, dtype=np.float64)).to(dev)
        assert fn.shape == (f.size(0), 3), "one target normal per face"
        vf = torch.from_numpy(np.ascontiguousarray(self.vta[0], dtype=np.int32)).to(dev)
        ni = torch.from_numpy(np.ascontiguousarray(self.vta[1], dtype=np.int32)).to(dev)
        out, scratch = torch.empty_like(v), torch.empty_like(v)
        _lib.check(_lib.load().ngpd_mesh_vertex_update(_lib.ptr(v), v.size(0), _lib.ptr(f), _lib.ptr(fn), _lib.ptr(vf), _lib.ptr(ni), int(k),
                                                       _lib.ptr(scratch), _lib.ptr(out), _lib.stream()), "ngpd_mesh_vertex_update")
        self.v[...] = out.cpu().numpy()
