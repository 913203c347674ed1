"""CPU-only checks: the C-ABI library loads and exports every symbol include/ngpd.h declares, the ctypes table
matches it, compute calls refuse to run without a device, and the host-side logic of the Python mirror."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ngpd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ngpd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    import ngpd_b200
    lib = ctypes.CDLL(ngpd_b200._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ngpd.h but not exported"
    assert set(names) == set(ngpd_b200._lib.SIGNATURES), set(names) ^ set(ngpd_b200._lib.SIGNATURES)
    assert lib.ngpd_version() == 100


def test_no_cpu_fallback():
    import ngpd_b200
    from ngpd_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    pc = ngpd_b200.Pointcloud(torch.rand(10, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ngpd_b200.Processor(pc)
    with pytest.raises(RuntimeError):
        _lib.Grid(torch.rand(10, 3))
    with pytest.raises(RuntimeError):
        ngpd_b200.TorchUtils.ChamferDistance(torch.rand(5, 3), torch.rand(5, 3))


def test_selection_container():
    from ngpd_b200 import Selection
    j = torch.tensor([0, 1, 2, 1, 0, 2, 2, 1, 0])
    s = Selection(torch.arange(3), j, torch.tensor([0, 3, 6, 9]))
    assert len(s) == 3 and s.uniform_k() == 3
    assert torch.equal(s[1], torch.tensor([1, 0, 2]))
    assert torch.equal(s.getEdgeIndex(), torch.stack([torch.arange(3).repeat_interleave(3), j]))
    f = s.filter(torch.tensor([2, 0]))
    assert torch.equal(f.i, torch.tensor([2, 0])) and torch.equal(f.j, torch.tensor([2, 1, 0, 0, 1, 2]))
    assert torch.equal(f.slices, torch.tensor([0, 3, 6]))
    src = torch.arange(9, dtype=torch.float32)
    assert torch.equal(s.scatter(src, "add"), torch.tensor([3.0, 12.0, 21.0]))
    assert torch.equal(s.scatter(src, "mean"), torch.tensor([1.0, 4.0, 7.0]))
    assert torch.equal(s.scatter(src, "max")[0], torch.tensor([2.0, 5.0, 8.0]))
    ragged = Selection(torch.arange(2), torch.tensor([1, 0, 1]), torch.tensor([0, 1, 3]))
    assert ragged.uniform_k() is None
    with pytest.raises(AssertionError):
        Selection(torch.arange(3), j.float(), torch.tensor([0, 3, 6, 9]))
    with pytest.raises(AssertionError):
        Selection(torch.arange(3), j, torch.tensor([1, 3, 6, 9]))


def test_range_boundaries():
    from ngpd_b200 import TorchUtils
    got = TorchUtils.rangeBoundariesToIndices(torch.tensor([5, 0, 9]), torch.tensor([8, 2, 9]))
    assert got.tolist() == [5, 6, 7, 0, 1]
    assert TorchUtils.rangeBoundariesToIndices(torch.tensor([3]), torch.tensor([3])).numel() == 0


def test_obj_roundtrip(tmp_path):
    from ngpd_b200 import Pointcloud
    v = torch.tensor([[0.0, 1.0, 2.0], [3.5, -4.25, 5.125], [1e-3, 2e5, -7.0]])
    n = torch.nn.functional.normalize(torch.tensor([[0.0, 0.0, 1.0], [1.0, 1.0, 0.0], [0.0, -2.0, 0.0]]), dim=1)
    p = tmp_path / "a.obj"
    Pointcloud(v, n).saveObj(str(p))
    back = Pointcloud.loadObj(str(p), device="cpu")
    assert torch.equal(back.v, v) and torch.allclose(back.n, n)
    with pytest.raises(FileExistsError):
        Pointcloud(v).saveObj(str(p))
    q = tmp_path / "mesh.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\nf 2//1 4//1 3//1\n")
    m = Pointcloud.loadObj(str(q), device="cpu")
    assert m.v.shape == (4, 3) and torch.allclose(m.n, torch.tensor([[0.0, 0.0, 1.0]]).expand(4, 3))
    with pytest.raises(AssertionError):
        Pointcloud(torch.zeros(3, 2))
    with pytest.raises(AssertionError):
        Pointcloud.loadObj(str(tmp_path / "missing.obj"))


def test_noise_validation_and_seeding():
    from ngpd_b200.GraphBuilder import Graph
    from ngpd_b200 import Noise
    g = Graph(pos=torch.zeros(50, 3), n=torch.tensor([[0.0, 0.0, 1.0]]).repeat(50, 1))
    nz = Noise(g)
    with pytest.raises(ValueError):
        nz.generateNoise(1.5, 1.0)
    torch.manual_seed(5)
    nz.generateNoise(0.3, 2.0, keepNormals=True)
    a = g.pos.clone()
    assert torch.equal(g.gt, torch.zeros(50, 3)) and (a[:, :2] == 0).all() and a[:, 2].std() > 0.2
    nz.resetNoise()
    torch.manual_seed(5)
    nz.generateNoise(0.3, 2.0, keepNormals=True)
    assert torch.equal(g.pos, a)


def test_acos_threshold_matches_oracle():
    import math
    import ngpd_oracle as O
    from ngpd_b200 import _lib
    for rho in (math.pi * 5 / 12, 0.9, 0.95, math.pi * 23 / 48):
        assert np.float32(_lib.acos_threshold(rho)) == O.acos_threshold(rho)


def test_sample_obj_properties(tmp_path):
    """Pointcloud.sampleObj (Object.py:135-156; torch_geometric's SamplePoints restated): points lie on the mesh, faces are
    drawn in proportion to their area, normals are the faces' unit normals, the global torch seed makes it reproducible."""
    import numpy as np
    import torch
    import ngpd_b200 as ng
    # a unit square in z = 0 (two triangles) and a 2 x 2 square in x = 3 (two triangles): area ratio 1 : 4
    obj = tmp_path / "two_squares.obj"
    obj.write_text("\n".join(["v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 3 0 0", "v 3 2 0", "v 3 2 2", "v 3 0 2",
                               "f 1 2 3", "f 1 3 4", "f 5 6 7", "f 5 7 8"]) + "\n")
    torch.manual_seed(0)
    pc = ng.Pointcloud.sampleObj(str(obj), 20000, device="cpu")
    v, n = pc.v.numpy(), pc.n.numpy()
    assert v.shape == (20000, 3) and n.shape == (20000, 3) and pc.hasNormals()
    small = np.abs(v[:, 2]) < 1e-6
    big = np.abs(v[:, 0] - 3) < 1e-6
    assert (small | big).all()
    assert (v[small, :2] >= -1e-6).all() and (v[small, :2] <= 1 + 1e-6).all()
    assert (v[big, 1:] >= -1e-6).all() and (v[big, 1:] <= 2 + 1e-6).all()
    assert abs(big.mean() - 0.8) < 0.02                                          # area 4 of 5
    assert np.allclose(np.abs(n[small]), [0, 0, 1], atol=1e-6) and np.allclose(np.abs(n[big]), [1, 0, 0], atol=1e-6)
    # uniform inside a face: the mean of the small square's samples is its centre
    assert np.abs(v[small, :2].mean(0) - 0.5).max() < 0.02
    torch.manual_seed(0)
    again = ng.Pointcloud.sampleObj(str(obj), 20000, device="cpu")
    assert torch.equal(again.v, pc.v) and torch.equal(again.n, pc.n)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_loader(tmp_path, fmt):
    """Pointcloud.loadPly (Object.py:119-132; the reference goes through Open3D): ascii and both binary layouts, extra vertex
    properties in between, a face element after the vertices, comments in the header."""
    import torch
    import ngpd_b200 as ng
    rng = np.random.default_rng(5)
    n = 257
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    red = rng.integers(0, 255, n).astype(np.uint8)
    nx = rng.normal(size=n)
    header = ["ply", f"format {fmt} 1.0", "comment made by a test", f"element vertex {n}", "property float x", "property uchar red",
              "property float y", "property double nx", "property float z", "element face 1", "property list uchar int vertex_indices",
              "end_header"]
    path = tmp_path / "cloud.ply"
    with open(path, "wb") as fh:
        fh.write(("\n".join(header) + "\n").encode("ascii"))
        if fmt == "ascii":
            for i in range(n):
                # repr of the fp32 value widened to a Python float round-trips exactly
                fh.write(f"{float(xyz[i, 0])!r} {int(red[i])} {float(xyz[i, 1])!r} {float(nx[i])!r} {float(xyz[i, 2])!r}\n".encode("ascii"))
            fh.write(b"3 0 1 2\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            rec = np.zeros(n, dtype=[("x", e + "f4"), ("red", "u1"), ("y", e + "f4"), ("nx", e + "f8"), ("z", e + "f4")])
            rec["x"], rec["red"], rec["y"], rec["nx"], rec["z"] = xyz[:, 0], red, xyz[:, 1], nx, xyz[:, 2]
            fh.write(rec.tobytes())
            fh.write(np.array([3], "u1").tobytes() + np.array([0, 1, 2], e + "i4").tobytes())
    pc = ng.Pointcloud.loadPly(str(path), device="cpu")
    assert pc.v.dtype == torch.float32 and tuple(pc.v.shape) == (n, 3) and pc.n is None
    assert np.array_equal(pc.v.numpy(), xyz)
    assert pc.file_path == str(path)
    with pytest.raises(AssertionError):
        ng.Pointcloud.loadPly(str(tmp_path / "cloud.obj"))


def test_better_vu_features():
    import torch
    import ngpd_b200 as ng
    w = torch.tensor([[0.0, 0.1, 0.9], [0.5, 0.6, 0.7], [0.0, 0.0, 0.01], [0.0, 0.2, 0.21]])
    dec = ng.Decomposition(w, torch.eye(3).repeat(4, 1, 1))
    l, k = 0.25, 6
    tau = 16.0 / k * l ** 2                                          # 0.1667
    assert torch.equal(dec.getBetterVUFeatures(l, k), (w < tau).sum(1) % 3)
    assert dec.getBetterVUFeatures(l, k).tolist() == [2, 0, 0, 1]
