"""GPU parity tests: the CUDA path (through the C ABI of libngpd.so and the Python mirror of the reference
classes) against the oracle on the same inputs and against golden vectors recorded from the unmodified reference.
Tolerances (north star): kNN indices and labels bit-exact; normals 1e-4 rad; positions 1e-5 relative; metrics 1e-6
relative.  fp32 throughout, fp64 only inside the kNN distance."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import ngpd_oracle as O
from conftest import angle_between

pytestmark = pytest.mark.gpu
RHO = math.pi * 5 / 12


@pytest.fixture(scope="module")
def ng():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ngpd_b200
    ngpd_b200._lib.load()
    return ngpd_b200


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def surface_cloud(n, seed=0, noise=0.004):
    """points on the unit cube's faces plus a torus around one edge: flat parts, creases, corners"""
    rng = np.random.default_rng(seed)
    nc = int(n * 0.6)
    face = rng.integers(0, 6, nc)
    uv = rng.uniform(-1, 1, (nc, 2))
    cube = np.zeros((nc, 3))
    for a in range(3):
        m = face // 2 == a
        o = [c for c in range(3) if c != a]
        cube[m, a] = np.where(face[m] % 2 == 0, -1.0, 1.0)
        cube[m, o[0]] = uv[m, 0]; cube[m, o[1]] = uv[m, 1]
    nt = n - nc
    u, v = rng.uniform(0, 2 * np.pi, nt), rng.uniform(0, 2 * np.pi, nt)
    tor = np.stack([(0.6 + 0.25 * np.cos(v)) * np.cos(u) + 1.0, (0.6 + 0.25 * np.cos(v)) * np.sin(u) + 1.0, 0.25 * np.sin(v)], 1)
    p = np.concatenate([cube, tor]) + rng.normal(0, noise, (n, 3))
    return rng.permutation(p).astype(np.float32)


# ------------------------------------------------------------------------------------------------------
# kNN
# ------------------------------------------------------------------------------------------------------
def test_knn_six_point_known_answer(ng):
    v = np.array([[-0.03785068, 0.12783747, 0.00448816], [-0.044779, 0.128887, 0.001905], [-0.06801, 0.151244, 0.037195],
                  [-0.070454, 0.150585, -0.043458], [-0.031026, 0.153728, -0.003546], [-0.040044, 0.15362, -0.008167]], dtype=np.float32)
    want = np.array([[1, 4, 5], [0, 5, 4], [1, 0, 5], [5, 4, 1], [5, 0, 1], [4, 1, 0]])
    g = ng._lib.Grid(cu(v), 4)
    assert np.array_equal(g.knn(cu(v), 4).cpu().numpy()[:, 1:], want)                   # scipy form: drop the self column
    assert np.array_equal(g.knn(cu(v), 3, ng._lib.KNN_SKIP_SELF).cpu().numpy(), want)  # torch_cluster form


@pytest.mark.parametrize("k,key", [(6, "knn6"), (16, "it0_knn16"), (8, "it0_knn8"), (16, "it1_knn16"), (8, "it1_knn8")])
def test_knn_fandisk_vs_reference(ng, fandisk, k, key):
    tree = fandisk["pos0"]
    query = fandisk["it1_pos_in"] if key.startswith("it1") else tree
    got = ng._lib.Grid(cu(tree), k).knn(cu(query), k, ng._lib.KNN_QUERY_IS_TREE).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, O.knn_bruteforce(tree, query, k))                        # bit-exact vs the oracle's rule
    assert O.tie_groups_equal(tree, query, got, fandisk[key].astype(np.int64))          # == SciPy up to exact ties


def test_knn_graph_noself_fandisk(ng, fandisk):
    tree = fandisk["pos0"]
    got = ng._lib.Grid(cu(tree), 12).knn(cu(tree), 12, ng._lib.KNN_SKIP_SELF | ng._lib.KNN_QUERY_IS_TREE).cpu().numpy()
    assert np.array_equal(got, O.knn_graph_noself(tree, 12))
    assert O.tie_groups_equal(tree, tree, got.astype(np.int64), fandisk["knn12_noself"].astype(np.int64))


@pytest.mark.parametrize("kind", ["volume", "surface", "lattice", "duplicates", "line", "clusters"])
@pytest.mark.parametrize("k", [1, 7, 16, 32, 64])
def test_knn_random_clouds_bit_exact(ng, kind, k):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{kind}{k}".encode()))
    n = 3000
    if kind == "volume":
        tree = rng.uniform(-1, 1, (n, 3))
    elif kind == "surface":
        tree = surface_cloud(n, 5)
    elif kind == "lattice":       # exact distance ties everywhere
        ax = np.arange(15, dtype=np.float64)
        tree = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)[:n] * 0.25
    elif kind == "duplicates":    # coincident points: ties at distance 0
        base = rng.uniform(0, 1, (n // 4, 3))
        tree = np.concatenate([base] * 4)
    elif kind == "line":          # degenerate bbox (two zero extents)
        tree = np.zeros((n, 3)); tree[:, 0] = rng.uniform(0, 5, n)
    else:                         # very uneven density + outliers far from everything
        tree = np.concatenate([rng.normal(0, 0.01, (n - 10, 3)), rng.uniform(50, 60, (10, 3))])
    tree = tree.astype(np.float32)
    query = np.concatenate([tree[rng.permutation(n)[:1500]] + rng.normal(0, 0.01, (1500, 3)).astype(np.float32),
                            rng.uniform(-3, 3, (200, 3)).astype(np.float32)]).astype(np.float32)   # some far outside the bbox
    g = ng._lib.Grid(cu(tree), k)
    idx, d2 = g.knn(cu(query), k, 0, with_d2=True)
    want = O.knn_bruteforce(tree, query, k)
    assert np.array_equal(idx.cpu().numpy(), want)
    ref_d2 = O.sqdist_rows_fp64(tree, query, want).astype(np.float32)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)


@pytest.mark.parametrize("k", [8, 16, 32, 64])
def test_knn_fast_path_equals_exact_search(ng, k):
    """the warp-lockstep fast path (+ fix-up list) returns exactly the rows of the exact shell search"""
    n = 600_000
    tree = cu(surface_cloud(n, 31, noise=0.002))
    query = tree + 0.001 * torch.randn_like(tree)
    g = ng._lib.Grid(tree, k)
    fast, d_fast = g.knn(query, k, ng._lib.KNN_QUERY_IS_TREE, with_d2=True)
    exact, d_exact = g.knn(query, k, ng._lib.KNN_QUERY_IS_TREE | ng._lib.KNN_EXACT_ONLY, with_d2=True)
    assert torch.equal(fast, exact) and torch.equal(d_fast, d_exact)
    # random (incoherent) query order and a foreign query set go through the ordering pass
    other = cu(surface_cloud(50_000, 32, noise=0.01))
    assert torch.equal(g.knn(other, k), g.knn(other, k, ng._lib.KNN_EXACT_ONLY))
    sess = ng._lib.Session(tree, k)
    sess.set_state(query, torch.nn.functional.normalize(torch.randn_like(tree), dim=1))
    s1, c1 = sess.mean_edge_length_parts(k)
    fix = sess.last_fixups()
    sess.set_knn_mode(1)
    s2, c2 = sess.mean_edge_length_parts(k)
    print(f"\nk={k}: {fix} of {n} queries ({fix / n:.3%}) went to the exact fix-up pass")
    assert c1 == c2 and abs(s1 - s2) <= 1e-9 * abs(s2)
    assert fix < 0.25 * n


def test_knn_fewer_points_than_k(ng):
    tree = np.random.default_rng(1).uniform(0, 1, (5, 3)).astype(np.float32)
    got = ng._lib.Grid(cu(tree), 8).knn(cu(tree), 8).cpu().numpy()
    want = O.knn_bruteforce(tree, tree, 8)
    assert np.array_equal(got, want) and (got[:, 5:] == 5).all()           # missing slots hold n, like SciPy
    one = ng._lib.Grid(cu(tree[:1]), 1).knn(cu(tree), 1).cpu().numpy()
    assert (one == 0).all()


def test_knn_bad_arguments(ng):
    tree = cu(np.zeros((4, 3), np.float32))
    g = ng._lib.Grid(tree, 4)
    with pytest.raises(ng._lib.NgpdError):
        g.knn(tree, 65)
    with pytest.raises(RuntimeError):
        g.knn(torch.zeros(4, 3), 2)                                       # CPU tensor: no fallback
    with pytest.raises(ng._lib.NgpdError):
        ng._lib.Grid(cu(np.full((4, 3), np.nan, np.float32)), 4)


def test_knn_large_properties(ng):
    """2 M points: rows sorted, self first, and a sample of rows equal to SciPy's KD-tree."""
    n, k = 2_000_000, 16
    tree = surface_cloud(n, 11, noise=0.0005)
    t = cu(tree)
    g = ng._lib.Grid(t, k)
    info = g.info()
    assert info.n == n and 2.0 < n / info.occupied_cells < 20.0
    idx, d2 = g.knn(t, k, ng._lib.KNN_QUERY_IS_TREE, with_d2=True)
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())
    assert bool((idx[:, 0] == torch.arange(n, device="cuda")).float().mean() > 0.9999)   # coincident points may swap
    assert bool((d2[:, 0] == 0).all())
    rows = np.random.default_rng(0).permutation(n)[:3000]
    want = O.knn_kdtree(tree, tree[rows], k)
    got = idx[cu(rows)].cpu().numpy().astype(np.int64)
    assert O.tie_groups_equal(tree, tree[rows], got, want)
    perm = g.order().cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(n))


# ------------------------------------------------------------------------------------------------------
# neighbourhood operators, teacher-forced on the reference's inputs for each stage
# ------------------------------------------------------------------------------------------------------
def test_pca_normals(ng, fandisk):
    pc = ng.Pointcloud(cu(fandisk["pos0"]))
    p = ng.Processor(pc)
    ei = p.graphBuilder.getKNNEdgeIndex(12)
    assert ei.shape == (2, len(fandisk["pos0"]) * 12) and ei.dtype == torch.long
    assert torch.equal(ei[0], torch.arange(len(fandisk["pos0"]), device="cuda").repeat_interleave(12))
    p.graph.edge_index = cu(np.stack([np.repeat(np.arange(len(fandisk["pos0"])), 12), fandisk["knn12_noself"].reshape(-1)]), torch.long)
    p.graphBuilder.setAndFlipNormals(flip=False)
    assert angle_between(p.graph.n.cpu().numpy(), fandisk["n_pca"]).max() < 1e-4      # sign-consistent with LAPACK
    vec = p.graphBuilder.getPVTDecompositionWithKNN(p.graph.edge_index)
    assert angle_between(vec[:, :, 0].cpu().numpy(), fandisk["n_pca"]).max() < 1e-4
    p.graphBuilder.flipNormals()
    agree = ((p.graph.n.cpu().numpy() * fandisk["n_flip"]).sum(1) > 0).mean()
    assert agree > 0.97, agree


def _processor(ng, pos, nrm, tree=None):
    pc = ng.Pointcloud(cu(tree if tree is not None else pos))
    p = ng.Processor(pc)
    p.graph.pos = cu(pos).clone()
    p.graph.n = cu(nrm).clone()
    return p


@pytest.mark.parametrize("it", [0, 1])
def test_nvt_smooth_classify_teacher_forced(ng, fandisk, it):
    t = f"it{it}_"
    p = _processor(ng, fandisk[t + "pos_in"], fandisk[t + "n_in"], tree=fandisk["pos0"])
    sel = p.selector.getKNNSelection(16)
    assert np.array_equal(sel.j.view(-1, 16).cpu().numpy(), O.knn_bruteforce(fandisk["pos0"], fandisk[t + "pos_in"], 16))
    # run the tensor stage on the reference's own neighbour table (differs from ours only on exact-tie rows)
    sel = ng.Selection(sel.i, cu(fandisk[t + "knn16"].reshape(-1), torch.long), sel.slices)
    n = len(sel)
    lib = ng._lib.load()
    pos, nrm = p.graph.pos, p.graph.n
    ev = torch.empty((n, 3), device="cuda"); vec = torch.empty((n, 3, 3), device="cuda"); T = torch.empty((n, 3, 3), device="cuda")
    sw = torch.empty(n, dtype=torch.int32, device="cuda")
    ng._lib.check(lib.ngpd_nvt(pos.data_ptr(), nrm.data_ptr(), sel.table().data_ptr(), None, None, n, 16, ng._lib.acos_threshold(RHO),
                               ev.data_ptr(), vec.data_ptr(), T.data_ptr(), sw.data_ptr(), None), "nvt")
    assert np.array_equal(T.cpu().numpy(), fandisk[t + "T1"])                         # voting tensors bit-exact
    assert np.abs(ev.cpu().numpy() - fandisk[t + "eigval1"]).max() < 1e-6
    dec = p.decompositionor.getBetterFilteredNVT(sel, nrm, RHO)
    assert torch.equal(dec.eigval, ev) and torch.equal(dec.eigvec, vec)
    gap = np.diff(fandisk[t + "eigval1"], axis=1).min(axis=1)
    signs = ((vec.cpu().numpy() * fandisk[t + "eigvec1"]).sum(1) > 0).all(1)
    assert signs[gap > 1e-2].all() and signs.mean() > 0.99                            # LAPACK's signs reproduced
    # smoothing given the reference's eigen-decomposition
    ref_dec = ng.Decomposition(cu(fandisk[t + "eigval1"]), cu(fandisk[t + "eigvec1"]))
    f = ref_dec.getVUSmoothedNormals(nrm)
    assert angle_between(f.cpu().numpy(), fandisk[t + "f_n"]).max() < 1e-4
    assert np.array_equal(f.cpu().numpy(), fandisk[t + "f_n"])
    # free-running smoothing: agreement at the reference's own noise floor
    f_own = dec.getVUSmoothedNormals(nrm)
    assert (angle_between(f_own.cpu().numpy(), fandisk[t + "f_n"]) > 1e-4).mean() < 0.0082
    # stage 2 on the reference's smoothed normals -> labels
    dec2 = p.decompositionor.getBetterFilteredNVT(sel, cu(fandisk[t + "f_n"]), RHO)
    assert np.abs(dec2.eigval.cpu().numpy() - fandisk[t + "eigval2"]).max() < 1e-6
    assert np.array_equal(dec2.getClasses().cpu().numpy(), fandisk[t + "classes"])    # labels bit-exact
    pla, lin, sph = dec2.getNVTFeatures()
    assert np.allclose(torch.stack([pla, lin, sph], 1).cpu().numpy(), fandisk[t + "features"], rtol=0, atol=2e-6)
    assert torch.equal(dec2.getVUFeatures(0.3), (dec2.eigval < 0.3).sum(1) % 3)


def test_k32_vs_reference(ng, fandisk, fandisk_k32):
    """BASELINE configs[2]'s neighbourhood size against the reference's Processor.getMyFeatureDecomposition(32): 32-NN rows
    bit-exact, voting tensors bit-exact on the reference's table, labels bit-exact given the reference's smoothed normals; then
    the fused session free-running (own search, own eigen-solver) for one iteration."""
    g = fandisk_k32
    p = _processor(ng, fandisk["pos0"], fandisk["n_flip"])
    sel = p.selector.getKNNSelection(32)
    ours = sel.j.view(-1, 32).cpu().numpy()
    assert np.array_equal(ours, O.knn_bruteforce(fandisk["pos0"], fandisk["pos0"], 32))
    assert O.tie_groups_equal(fandisk["pos0"], fandisk["pos0"], ours, g["knn32"].astype(np.int64))
    sel = ng.Selection(sel.i, cu(g["knn32"].reshape(-1), torch.long), sel.slices)
    n = len(sel)
    lib = ng._lib.load()
    ev = torch.empty((n, 3), device="cuda"); vec = torch.empty((n, 3, 3), device="cuda"); T = torch.empty((n, 3, 3), device="cuda")
    sw = torch.empty(n, dtype=torch.int32, device="cuda")
    ng._lib.check(lib.ngpd_nvt(p.graph.pos.data_ptr(), p.graph.n.data_ptr(), sel.table().data_ptr(), None, None, n, 32,
                               ng._lib.acos_threshold(RHO), ev.data_ptr(), vec.data_ptr(), T.data_ptr(), sw.data_ptr(), None), "nvt")
    assert np.array_equal(T.cpu().numpy(), g["T1"])
    assert np.abs(ev.cpu().numpy() - g["eigval1"]).max() < 1e-6
    dec2 = p.decompositionor.getBetterFilteredNVT(sel, cu(g["f_n"]), RHO)
    assert np.abs(dec2.eigval.cpu().numpy() - g["eigval2"]).max() < 1e-6
    assert np.array_equal(dec2.getClasses().cpu().numpy(), g["classes"])
    # fused session, k_feature = 32 (re-ranking tier with 64 candidates from the second iteration on)
    sess = ng._lib.Session(cu(fandisk["pos0"]), 32)
    sess.set_state(cu(fandisk["pos0"]), cu(fandisk["n_flip"]))
    sess.step(ng._lib.make_params(k_feature=32, dmax=float(np.float32(2) * g["l"])))
    pos, fn, lab = sess.get_state(True)
    differ = int((lab.cpu().numpy() != g["classes"]).sum())
    bad_n = float((angle_between(fn.cpu().numpy(), g["f_n"]) > 1e-4).mean())
    err = np.abs(pos.cpu().numpy() - g["pos_after"]).max(axis=1) / np.abs(fandisk["pos0"]).max()
    print(f"\nk=32 free-running: {differ} labels differ, normals > 1e-4 rad: {bad_n:.4%}, positions > 1e-5: {(err > 1e-5).mean():.4%}")
    # the host build of the same per-point math differs from the reference in 1 label here (tests/test_hostmath.py); the k = 16
    # yardsticks of test_session_labels_vs_reference bound the rest
    assert differ <= 7
    assert bad_n < 0.0080
    assert (err > 1e-5).mean() < 0.0377


@pytest.mark.parametrize("it", [0, 1])
def test_update_steps_teacher_forced(ng, fandisk, it):
    t = f"it{it}_"
    p = _processor(ng, fandisk[t + "pos_in"], fandisk[t + "n_in"], tree=fandisk["pos0"])
    f = cu(fandisk[t + "f_n"]); cls = cu(fandisk[t + "classes"]).long()
    edge = cu(np.ascontiguousarray(fandisk[t + "eigvec2"][:, :, 0]))
    n = len(fandisk["pos0"])
    sel8 = ng.Selection(torch.arange(n, device="cuda"), cu(fandisk[t + "knn8"].reshape(-1), torch.long), torch.arange(n + 1, device="cuda") * 8)
    d = float(np.float32(2) * fandisk["l"])
    scale = np.abs(fandisk[t + "pos_in"]).max()
    for key, alpha in ((0, 1), (1, 0.2), (2, 1)):
        rows = (cls == key).nonzero().flatten()
        sub = sel8.filter(rows)
        if key == 0:
            new = p.denoiser.flat_step(sub, f, d, alpha)
        elif key == 1:
            new = p.denoiser.edge_step(sub, f, edge, d, alpha)
        else:
            new = p.denoiser.feature_step(sub, f, d, alpha)
        ref = fandisk[t + f"pos_after_class{key}"]
        assert np.abs(new.cpu().numpy() - ref[rows.cpu().numpy()]).max() / scale < 1e-5, key
        p.graph.pos = cu(ref).clone(); p.denoiser.graph = p.graph
    # the remaining step kinds on all points (recorded after iteration 0's class loop)
    if it == 0:
        allsel = sel8
        feat = p.denoiser.feature_step(allsel, f, d, 0.5)
        assert np.abs(feat.cpu().numpy() - fandisk["feature_all"]).max() / scale < 1e-5
        cor = p.denoiser.corner_step(allsel, f, d, 0.1).cpu().numpy()
        nj = fandisk[t + "f_n"][fandisk[t + "knn8"]]
        ok = O.solve_condition(np.einsum("nki,nkj->nij", nj, nj)) < 1e3
        assert np.abs(cor - fandisk["corner_all"])[ok].max() / scale < 1e-5
        assert torch.equal(p.denoiser.dummy_step(allsel, f, d), p.graph.pos)


def test_ragged_selection_rows(ng, fandisk):
    """CSR rows of different length (what a radius selection produces) go through the same kernels."""
    pos, nrm = fandisk["it0_pos_in"], fandisk["it0_n_in"]
    p = _processor(ng, pos, nrm)
    n = 500
    rng = np.random.default_rng(2)
    lens = rng.integers(3, 17, n)
    full = fandisk["it0_knn16"][:n]
    j = np.concatenate([full[r, :lens[r]] for r in range(n)])
    slices = np.concatenate([[0], np.cumsum(lens)])
    sel = ng.Selection(torch.arange(n, device="cuda"), cu(j, torch.long), cu(slices, torch.long))
    dec = p.decompositionor.getBetterFilteredNVT(sel, p.graph.n, RHO)
    xt = O.acos_threshold(RHO)
    for r in (0, 17, 499):
        w, V, T, _ = O.nvt(pos, nrm, np.array([r]), full[r:r + 1, :lens[r]], xt)
        assert np.abs(dec.eigval[r].cpu().numpy() - w[0]).max() < 1e-6


# ------------------------------------------------------------------------------------------------------
# metrics
# ------------------------------------------------------------------------------------------------------
def test_metrics_vs_reference(ng, fandisk):
    gt, pos = cu(fandisk["gt"]), cu(fandisk["pos_final"])
    T = ng.TorchUtils
    cd = T.ChamferDistance(gt, pos).cpu().numpy()
    ref = fandisk["cd_final"]
    assert cd.shape == ref.shape
    assert abs(cd.mean(dtype=np.float64) - ref.mean(dtype=np.float64)) / ref.mean(dtype=np.float64) < 1e-6
    assert np.allclose(cd, ref, rtol=1e-5, atol=1e-9)
    assert np.allclose(T.PaperDistance(gt, pos).cpu().numpy(), fandisk["paper_final"], rtol=1e-5, atol=1e-9)
    assert np.allclose(T.HausdorffDistance(gt, pos).cpu().numpy(), fandisk["hausdorff_final"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(T.SingleChamferDistance(gt, pos).cpu().numpy(), cd[:len(fandisk["pos_final"])])
    assert abs(float(T.pointcloudRadius(pos)) - float(fandisk["radius_final"])) / float(fandisk["radius_final"]) < 1e-6
    cd0 = T.ChamferDistance(gt, cu(fandisk["pos0"])).cpu().numpy()
    assert abs(cd0.mean(dtype=np.float64) - fandisk["cd_initial"].mean(dtype=np.float64)) / fandisk["cd_initial"].mean(dtype=np.float64) < 1e-6


def test_chamfer_properties_large(ng):
    a = cu(surface_cloud(1_000_000, 3))
    T = ng.TorchUtils
    assert float(T.ChamferDistance(a, a).abs().max()) == 0.0                 # identity
    shift = a + torch.tensor([0.0, 0.0, 5.0], device="cuda")
    cd = T.ChamferDistance(a, shift)
    assert float(cd.min()) >= 0 and cd.numel() == 2 * a.size(0)
    b = a[torch.randperm(a.size(0), device="cuda")[:200_000]]
    one = T.ChamferDistance(a, b)
    assert float(one[:b.size(0)].max()) == 0.0                               # subset -> superset distances vanish
    m1, m2 = float(T.ChamferDistance(a, b).double().mean()), float(T.ChamferDistance(b, a).double().mean())
    assert abs(m1 - m2) / m1 < 1e-9                                          # mean over the concatenation is symmetric


# ------------------------------------------------------------------------------------------------------
# the loops
# ------------------------------------------------------------------------------------------------------
def _fandisk_processor(ng, fandisk):
    p = ng.Processor(ng.Pointcloud(cu(fandisk["pos0"]).clone()))
    p.graph.n = cu(fandisk["n_flip"]).clone()
    return p


def denoise_unfused(ng, p, k_feature: int = 16, k_update: int = 8, iterations: int = 2):
    """Processor.denoise written against the public operators, call for call as the reference does it (Processor.py:119-139):
    the cross-check of the fused session (test scaffolding; it lived in the product class in round 1)."""
    g = p.graph
    l = ng.TorchUtils.averageEdgeLength(g.pos, p.selector.getKNNSelection(6).getEdgeIndex())
    d = float(2 * l)
    alphas = [1, 0.2, 1]
    for _ in range(iterations):
        decomposition, f_n = p.getMyFeatureDecomposition(k_feature)
        classes = decomposition.getClasses()
        selection = p.selector.getKNNSelection(k_update)
        for key in range(3):
            indices = (classes == key).nonzero().flatten()
            if indices.size(0) == 0:
                continue
            if key == 0:
                new_pos = p.denoiser.flat_step(selection.filter(indices), f_n, d, alphas[key])
            elif key == 1:
                new_pos = p.denoiser.edge_step(selection.filter(indices), f_n, decomposition.eigvec[..., 0], d, alphas[key])
            else:
                new_pos = p.denoiser.feature_step(selection.filter(indices), f_n, d, alphas[key])
            g.pos[indices] = new_pos
        g.n = f_n


def test_denoise_fused_equals_unfused(ng, fandisk):
    a = _fandisk_processor(ng, fandisk); a.denoise()
    b = _fandisk_processor(ng, fandisk); denoise_unfused(ng, b)
    scale = float(a.graph.pos.abs().max())
    assert float((a.graph.pos - b.graph.pos).abs().max()) / scale < 2e-6
    assert angle_between(a.graph.n.cpu().numpy(), b.graph.n.cpu().numpy()).max() < 1e-4


@pytest.mark.parametrize("k_f,k_u", [(32, 8), (8, 6), (64, 16), (12, 5)])
def test_denoise_fused_equals_unfused_other_neighbourhoods(ng, fandisk, k_f, k_u):
    """The fused session against the per-stage public operators for the neighbourhood sizes of BASELINE configs[2] (k = 32),
    the notebooks (64) and sizes without a specialised kernel (12 / 5): same labels' effect on positions, same normals."""
    a = _fandisk_processor(ng, fandisk); a._denoise_fused(k_f, k_u, 2)
    b = _fandisk_processor(ng, fandisk); denoise_unfused(ng, b, k_f, k_u, 2)
    scale = float(a.graph.pos.abs().max())
    assert float((a.graph.pos - b.graph.pos).abs().max()) / scale < 2e-6
    assert angle_between(a.graph.n.cpu().numpy(), b.graph.n.cpu().numpy()).max() < 1e-4


def test_denoise_vs_reference_free_running(ng, fandisk):
    """Processor.denoise() end to end against the reference's result.  The reference's smoothing depends on LAPACK's
    eigenvector signs inside rank-deficient tensors (a 1-ulp change of its own input moves 0.82 % of its normals by
    > 1e-4 rad and flips 10/6475 labels: SURVEY.md 8a row 5), so the free-running comparison is statistical, held to
    that noise floor; the stage-wise tests above carry the strict tolerances."""
    p = _fandisk_processor(ng, fandisk)
    pos_before = p.graph.pos
    p.denoise()
    assert p.graph.pos is pos_before                                          # updated in place, like the reference
    pos, nrm = p.graph.pos.cpu().numpy(), p.graph.n.cpu().numpy()
    scale = np.abs(fandisk["pos_final"]).max()
    err = np.abs(pos - fandisk["pos_final"]).max(axis=1) / scale
    ang = angle_between(nrm, fandisk["n_final"])
    print(f"\nfree-running denoise(): positions >1e-5: {(err > 1e-5).mean():.4%} (max {err.max():.2e}); normals >1e-4 rad: {(ang > 1e-4).mean():.4%}")
    # yardstick: the reference run twice, the second time with its input normals moved by 1 ulp, differs from itself by
    # 12.57 % of positions (> 1e-5, max 4.2e-3), 6.75 % of normals (> 1e-4 rad) and 20 labels after these two iterations
    # (measured with the oracle, same LAPACK; DESIGN.md "parity protocol")
    assert (err > 1e-5).mean() < 0.1257 and err.max() < 1e-2
    assert (ang > 1e-4).mean() < 0.0675
    cd = ng.TorchUtils.ChamferDistance(cu(fandisk["gt"]), p.graph.pos).double().mean().item()
    ref = fandisk["cd_final"].mean(dtype=np.float64)
    # Chamfer yardstick (tests/golden/noise_floor.py, three 1-ulp trials): the reference differs from itself by 3.2e-4 .. 8.5e-4 relative
    assert abs(cd - ref) / ref < 3e-3


def test_session_labels_vs_reference(ng, fandisk):
    """Labels of iteration 0 from the fused session (own kNN, own eigen-solver) vs the reference."""
    sess = ng._lib.Session(cu(fandisk["pos0"]), 16)
    sess.set_state(cu(fandisk["pos0"]), cu(fandisk["n_flip"]))
    d = float(np.float32(2) * fandisk["l"])
    s, c = sess.mean_edge_length_parts(6)
    assert abs(s / c - float(fandisk["l"])) / float(fandisk["l"]) < 1e-6
    sess.step(ng._lib.make_params(dmax=d))
    pos, fn, lab = sess.get_state(True)
    agree = (lab.cpu().numpy() == fandisk["it0_classes"]).mean()
    print(f"\nfree-running labels agree with the reference on {agree:.4%} of points")
    # yardstick after ONE iteration (reference vs itself under a 1-ulp change of the input normals): 7 labels,
    # 0.80 % of normals, 3.77 % of positions
    assert (lab.cpu().numpy() != fandisk["it0_classes"]).sum() <= 7
    assert (angle_between(fn.cpu().numpy(), fandisk["it0_f_n"]) > 1e-4).mean() < 0.0080
    err = np.abs(pos.cpu().numpy() - fandisk["it0_pos_after_class2"]).max(axis=1) / np.abs(fandisk["pos0"]).max()
    assert (err > 1e-5).mean() < 0.0377
    assert 8 <= sess.launch_count() <= 16


def test_until_minimum_error_loop(ng, until_min):
    """BASELINE config 2 on the recorded cloud: same stopping iteration and Chamfer history as the reference."""
    p = ng.Processor(ng.Pointcloud(cu(until_min["pos0"]).clone()))
    p.graph.n = cu(until_min["n_flip"]).clone()
    l = float(until_min["l"])
    hist = []

    def cd(a, b):
        r = ng.TorchUtils.ChamferDistance(a, b)
        hist.append(float(r.double().mean()))
        return r

    strategy = {0: p.denoiser.flat_step, 1: p.denoiser.feature_step, 2: p.denoiser.feature_step}
    noisy = p.graph.pos.clone()
    best, prev_err, iters = p.denoiseUntilMinimumError(cu(until_min["gt"]), strategy, k=8, alpha=[1, 0.2, 1], d=2 * l, error_funcs=[cd])
    ref_hist = until_min["cd_history"]
    print("\nCD history ours", hist, "\nCD history ref ", ref_hist.tolist())
    assert iters == int(until_min["iterations"])
    assert len(hist) == len(ref_hist)
    assert np.allclose(hist[:2], ref_hist[:2], rtol=1e-6)
    assert np.allclose(hist, ref_hist, rtol=2e-3)
    assert torch.equal(p.graph.pos, noisy)                                   # reset to the noisy input on exit
    scale = np.abs(until_min["pos0"]).max()
    err = np.abs(best.cpu().numpy() - until_min["pos_returned"]).max(axis=1) / scale
    # yardstick (tests/golden/noise_floor.py): the reference's algorithm differs from ITSELF on 23.9 - 24.5 % of the positions
    # (> 1e-5, max 2.5e-3) after these three iterations when its input normals move by 1 ulp
    print(f"returned positions >1e-5 from the reference's: {(err > 1e-5).mean():.4%} (max {err.max():.2e})")
    assert (err > 1e-5).mean() < 0.239 and err.max() < 1e-2


def test_generic_strategy_path(ng, fandisk):
    """A strategy holding a user callable falls back to the operator-by-operator loop."""
    p = _fandisk_processor(ng, fandisk)
    calls = []

    def my_step(selection, n, d, alpha):
        calls.append(len(selection))
        return p.denoiser.feature_step(selection, n, d, alpha)

    gt = cu(fandisk["gt"])
    strategy = {0: p.denoiser.flat_step, 1: my_step, 2: p.denoiser.dummy_step}
    best, err, it = p.denoiseUntilMinimumError(gt, strategy, k=8, alpha=[1, 0.2, 1], d=float(2 * fandisk["l"]),
                                              error_funcs=[ng.TorchUtils.ChamferDistance])
    assert calls and it >= 0


def test_cube_labels_vs_reference(ng, cube):
    p = ng.Processor(ng.Pointcloud(cu(cube["pos_clean"])))
    p.graph.pos = cu(cube["pos"]).clone(); p.graph.n = cu(cube["n_flip"]).clone()
    p.decompositionor.graph = p.graph; p.selector.graph = p.graph
    sel = p.selector.getKNNSelection(16)
    assert O.tie_groups_equal(cube["pos_clean"], cube["pos"], sel.j.view(-1, 16).cpu().numpy(), cube["knn16"].astype(np.int64))
    sel = ng.Selection(sel.i, cu(cube["knn16"].reshape(-1), torch.long), sel.slices)       # reference's tie order
    ang = float(cube["angle"])
    nvt = p.decompositionor.getBetterFilteredNVT(sel, p.graph.n, ang)
    f = nvt.getVUSmoothedNormals(p.graph.n)
    dec = p.decompositionor.getBetterFilteredNVT(sel, f, ang)
    agree = (dec.getClasses().cpu().numpy() == cube["classes"]).mean()
    assert agree > 0.97, agree
    # FeatureFix.ipynb#c1 rule: label = (#coordinates with |x| = 1) - 1; the reference scores 91.7 % on this cloud
    assert (dec.getClasses().cpu().numpy() == cube["gt_label"]).mean() > 0.88


def test_preprocess_pointcloud(ng, fandisk):
    torch.manual_seed(0)
    p = ng.Processor(ng.Pointcloud(cu(fandisk["gt"]).clone()))
    p.preprocessPointcloud(k=12, noise_level=0.3)
    g = p.graph
    assert torch.equal(g.gt, cu(fandisk["gt"])) and g.n.shape == g.pos.shape
    disp = (g.pos - g.gt).norm(dim=1)
    l = float(ng.TorchUtils.averageEdgeLength(g.gt, g.edge_index))
    assert 0.25 * l < float(disp.square().mean().sqrt()) < 0.35 * l             # rms displacement = sigma = 0.3 * mean edge length
    assert torch.allclose(g.n.norm(dim=1), torch.ones(g.num_nodes, device="cuda"), atol=1e-5)
    # selector stays frozen on the clean cloud
    assert torch.equal(p.selector.tree_pos, cu(fandisk["gt"]))


def test_session_large_runs_and_is_deterministic(ng):
    n = 1_000_000
    cloud = cu(surface_cloud(n, 21, noise=0.001))
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
    outs = []
    for _ in range(2):
        sess = ng._lib.Session(cloud, 16)
        sess.set_state(cloud, nrm)
        s, c = sess.mean_edge_length_parts(6)
        params = ng._lib.make_params(dmax=2 * s / c)
        sess.step(params); sess.step(params)
        outs.append(sess.get_state(True))
    for a, b in zip(*outs):
        assert torch.equal(a, b)                                              # no atomics in any per-point result
    # the three kNN modes (re-ranking + streaming tiers, exact shell search, streaming tiers only) give the same run
    for mode in (1, 2):
        sess = ng._lib.Session(cloud, 16)
        sess.set_state(cloud, nrm)
        sess.set_knn_mode(mode)
        sess.step(params); sess.step(params)
        for a, b in zip(outs[0], sess.get_state(True)):
            assert torch.equal(a, b)
    pos, fn, lab = outs[0]
    assert torch.isfinite(pos).all() and torch.isfinite(fn).all()
    assert torch.allclose(fn.norm(dim=1), torch.ones(n, device="cuda"), atol=1e-4)
    assert int(lab.max()) <= 2


def _session_table(ng, sess, k):
    """the session's neighbour table (tree positions, tree order) as a tensor"""
    import ctypes
    base = ng._lib.load().ngpd_session_buffer(sess._h, 6)
    out = torch.empty((sess.n, k), dtype=torch.int32, device="cuda")
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(base), ctypes.c_size_t(4 * sess.n * k), 3)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("k_f,k_u", [(16, 8), (8, 8), (32, 8)])
def test_knn_rerank_tier_is_exact(ng, k_f, k_u):
    """Temporal coherence: after the first search most rows are answered by re-ranking the stored candidates (tier 0).
    Every iteration's neighbour table must equal the exact shell search's, also after positions were replaced from
    outside (set_state) with some points thrown far away."""
    n = 400_000
    cloud = cu(surface_cloud(n, 77, noise=0.002))
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
    a, b = ng._lib.Session(cloud, k_f), ng._lib.Session(cloud, k_f)
    b.set_knn_mode(1)
    for sess in (a, b):
        sess.set_state(cloud, nrm)
    s, c = a.mean_edge_length_parts(6)
    params = ng._lib.make_params(k_feature=k_f, k_update=k_u, dmax=2 * s / c)
    answered = []
    for it in range(5):
        if it == 3:
            # outside edit: everything jitters a little, 1 % of the points jump by many cells
            pos, nr, _ = a.get_state(False)
            g = torch.Generator(device="cuda").manual_seed(5)
            pos = pos + 0.05 * (s / c) * torch.randn(pos.shape, device="cuda", generator=g)
            far = torch.rand(n, device="cuda", generator=g) < 0.01
            pos[far] += 0.2 * torch.randn((int(far.sum()), 3), device="cuda", generator=g)
            for sess in (a, b):
                sess.set_state(pos, nr)
        a.step(params); b.step(params)
        t0, t1, t2 = a.knn_stats()
        answered.append(1.0 - t0 / n if it > 0 else 0.0)
        assert torch.equal(_session_table(ng, a, k_f), _session_table(ng, b, k_f)), f"iteration {it}"
        # the mean-edge-length pass (shorter rows, untracked) reuses the candidates as well
        assert a.mean_edge_length_parts(6) == b.mean_edge_length_parts(6)
    for x, y in zip(a.get_state(True), b.get_state(True)):
        assert torch.equal(x, y)
    print(f"\nk_f={k_f}: share of rows answered by the re-ranking tier per iteration: {[round(v, 4) for v in answered]}")
    assert answered[1] > 0.5 and answered[4] > 0.5


def test_eigh3_device_equals_host_transcription(ng):
    """ngpd_eigh3 on the GPU (n = 3 specialisation, branch-free correctly rounded division / square root inside the
    safe exponent window, plain operators outside it) is bit-identical to the array-indexed LAPACK transcription compiled
    for the host (tests/hostmath/eig3_generic.h)."""
    import ctypes
    import subprocess
    from conftest import ROOT
    hmdir = os.path.join(ROOT, "tests", "hostmath")
    so = os.path.join(hmdir, "libhostmath.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", os.path.join(hmdir, "hostmath.cpp"), "-o", so], check=True)
    hm = ctypes.CDLL(so)
    rng = np.random.default_rng(5)
    m = 300_000

    def sym(a):
        return (a + a.transpose(0, 2, 1)) / 2

    def votes(k, spread):
        base = rng.normal(size=(m, 1, 3)); base /= np.linalg.norm(base, axis=2, keepdims=True)
        n = base + spread * rng.normal(size=(m, k, 3)); n /= np.linalg.norm(n, axis=2, keepdims=True)
        n = n.astype(np.float32)
        return np.einsum("mki,mkj->mij", n, n) / np.float32(k)

    cases = {"general": sym(rng.normal(size=(m, 3, 3))), "votes_flat": votes(16, 0.05), "votes_noisy": votes(16, 0.5),
             "votes_crease": (votes(8, 0.02) + votes(8, 0.02)) / 2, "votes_rank1": votes(1, 0.0), "votes_rank2": votes(2, 0.3),
             "zero": np.zeros((4, 3, 3)), "identity": np.tile(np.eye(3), (4, 1, 1)),
             "window_low": sym(rng.normal(size=(m, 3, 3))) * 2.0 ** -9, "window_high": sym(rng.normal(size=(m, 3, 3))) * 2.0 ** 7,
             "tiny": sym(rng.normal(size=(m, 3, 3))) * 1e-20, "huge": sym(rng.normal(size=(m, 3, 3))) * 1e15}
    ax = np.eye(3)[rng.integers(0, 3, (m, 16))] * rng.choice([-1, 1], (m, 16, 1))
    cases["axis_aligned"] = np.einsum("mki,mkj->mij", ax, ax) / 16
    mixed = sym(rng.normal(size=(m, 3, 3))); mixed[:, 1:, 1:] *= 1e-12; mixed[:, 0, 1:] *= 1e-7; mixed[:, 1:, 0] *= 1e-7
    cases["mixed_scales"] = mixed
    pts = rng.normal(size=(m, 12, 3)) * np.array([1, 1, 0.01]); pts -= pts.mean(1, keepdims=True)
    cases["covariance"] = np.einsum("mki,mkj->mij", pts, pts)
    lib = ng._lib.load()
    for name, T in cases.items():
        T = np.ascontiguousarray(T.astype(np.float32)); n = len(T)
        w1 = np.zeros((n, 3), np.float32); V1 = np.zeros((n, 3, 3), np.float32)
        hm.hm_eigh3_generic(T.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), w1.ctypes.data_as(ctypes.c_void_p), V1.ctypes.data_as(ctypes.c_void_p))
        Td = cu(T); wd = torch.empty((n, 3), device="cuda"); Vd = torch.empty((n, 3, 3), device="cuda")
        ng._lib.check(lib.ngpd_eigh3(Td.data_ptr(), n, wd.data_ptr(), Vd.data_ptr(), None), "eigh3")
        w2, V2 = wd.cpu().numpy(), Vd.cpu().numpy()
        bad = ~((w1.view(np.uint32) == w2.view(np.uint32)).all(1) & (V1.view(np.uint32) == V2.view(np.uint32)).all((1, 2)))
        print(f"\n{name}: {bad.sum()} of {n} tensors differ from the host transcription")
        assert bad.sum() == 0, name


# ------------------------------------------------------------------------------------------------------
# host-buffer entry points (the end-to-end path of bench.py) and the phase-by-phase driver
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("iterations", [1, 2])
def test_run_host_equals_device_steps(ng, iterations):
    """ngpd_session_run_host (pinned host buffers, transfers overlapped with the search and the updates on a second
    stream) returns exactly what set_state + step + get_state give; calling it again continues from the new state."""
    n = 300_000
    cloud = surface_cloud(n, 33, noise=0.002)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(3)), dim=1)
    a, b = ng._lib.Session(cu(cloud), 16), ng._lib.Session(cu(cloud), 16)
    a.set_state(cu(cloud), nrm.cuda())
    s, c = a.mean_edge_length_parts(6)
    params = ng._lib.make_params(dmax=2 * s / c)
    pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
    pos_h, nrm_h, pos_o, nrm_o, lab_o = pin(n, 3), pin(n, 3), pin(n, 3), pin(n, 3), pin(n, dtype=torch.uint8)
    pos_h.copy_(torch.from_numpy(cloud)); nrm_h.copy_(nrm)
    for call in range(2):
        for _ in range(iterations):
            a.step(params)
        want = [t.cpu() for t in a.get_state(True)]
        b.run_host(params, iterations, pos_h, nrm_h, pos_o, nrm_o, lab_o)
        for w, g, name in zip(want, (pos_o, nrm_o, lab_o), ("positions", "normals", "labels")):
            assert torch.equal(w, g), f"call {call}: {name}"
        pos_h.copy_(pos_o); nrm_h.copy_(nrm_o)
    # outputs are optional
    b.run_host(params, 1, pos_h, nrm_h, pos_o, None, None)
    a.step(params)
    assert torch.equal(a.get_state(False)[0].cpu(), pos_o)


def test_denoise_host_one_shot(ng):
    import ctypes
    n = 50_000
    cloud = surface_cloud(n, 34, noise=0.003)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=torch.Generator().manual_seed(4)), dim=1).numpy()
    sess = ng._lib.Session(cu(cloud), 16)
    sess.set_state(cu(cloud), cu(nrm))
    s, c = sess.mean_edge_length_parts(6)
    params = ng._lib.make_params(dmax=2 * s / c)
    sess.step(params); sess.step(params)
    want = [t.cpu().numpy() for t in sess.get_state(True)]
    pos_o, nrm_o, lab_o = np.empty_like(cloud), np.empty_like(nrm), np.empty(n, np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = ng._lib.load().ngpd_denoise_host(vp(cloud), vp(cloud), vp(nrm), n, ctypes.byref(params), 2, vp(pos_o), vp(nrm_o), vp(lab_o))
    assert rc == 0, ng._lib.load().ngpd_last_error()
    for w, g in zip(want, (pos_o, nrm_o, lab_o)):
        assert np.array_equal(w, g)


def test_phase_driver_equals_fused_step(ng):
    """The phase entry points a multi-GPU driver calls one by one (with the flat-step scalars recomputed by their own
    kernels after an outside edit of the positions) give the same iteration as ngpd_session_step, whose stage-2 kernel
    produces the class-0 sums and the minority-class row lists itself."""
    import ctypes
    n = 200_000
    cloud = cu(surface_cloud(n, 35, noise=0.002))
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
    lib = ng._lib.load()
    outs = []
    for mode in range(3):
        sess = ng._lib.Session(cloud, 16)
        sess.set_state(cloud, nrm)
        s, c = sess.mean_edge_length_parts(6)
        for strategy in ((ng._lib.STEP_FLAT, ng._lib.STEP_EDGE, ng._lib.STEP_FEATURE), (ng._lib.STEP_FEATURE, ng._lib.STEP_FLAT, ng._lib.STEP_CORNER)):
            params = ng._lib.make_params(dmax=2 * s / c, strategy=strategy)
            ref = ctypes.byref(params)
            if mode == 0:
                sess.step(params)
                continue
            st = ng._lib.stream
            assert lib.ngpd_session_phase_features(sess._h, ref, 0, st()) == 0
            assert lib.ngpd_session_phase_features(sess._h, ref, 1, st()) == 0
            if mode == 2:
                # a no-op "halo import" of positions: drops the fused sums, the stand-alone reduction must give the same centre
                rows = torch.arange(4, dtype=torch.int32, device="cuda")
                buf = torch.empty((4, 4), device="cuda")
                assert lib.ngpd_session_export_rows(sess._h, 0, rows.data_ptr(), 4, buf.data_ptr(), st()) == 0
                assert lib.ngpd_session_import_rows(sess._h, 0, rows.data_ptr(), 4, buf.data_ptr(), st()) == 0
            for key in range(3):
                if params.strategy[key] == ng._lib.STEP_FLAT:
                    assert lib.ngpd_session_phase_flat_scalars(sess._h, ref, key, 0, st()) == 0
                    assert lib.ngpd_session_phase_flat_scalars(sess._h, ref, key, 1, st()) == 0
                assert lib.ngpd_session_phase_update(sess._h, ref, key, st()) == 0
            assert lib.ngpd_session_phase_commit_normals(sess._h) == 0
            if mode == 2:
                break                                                         # one iteration is enough for the comparison below
        outs.append(sess.get_state(True))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    # mode 2 sums the class in a different order (atomics instead of the fixed per-block order): the centre may differ in
    # the last bit, the positions by a few ulp.  Compared after the first iteration (labels are decided before anything moves).
    sess = ng._lib.Session(cloud, 16)
    sess.set_state(cloud, nrm)
    sess.step(ng._lib.make_params(dmax=2 * s / c))
    one = sess.get_state(True)
    assert torch.equal(one[2], outs[2][2])
    assert (one[0] - outs[2][0]).abs().max().item() <= 1e-6 * one[0].abs().max().item()


# ------------------------------------------------------------------------------------------------------
# Yadav-2018 baseline path ("CPSD"): radius selection, normal-filtered NVT / PVT, VU labels, the
# notebook's loop -- against vectors recorded from the reference (tests/golden/make_golden_cpsd.py)
# ------------------------------------------------------------------------------------------------------
def _cpsd_processor(ng, cpsd, it):
    """Processor whose tree is the noisy input and whose current state is the reference's input of iteration `it`"""
    return _processor(ng, cpsd[f"it{it}_pos_in"], cpsd[f"it{it}_n_in"], tree=cpsd["pos0"])


@pytest.mark.parametrize("it", [0, 1, 2])
def test_cpsd_ball_selection(ng, cpsd, it):
    p = _cpsd_processor(ng, cpsd, it)
    sel = p.selector.getPointsInRangeSelection(float(cpsd["d"]))
    assert torch.equal(sel.slices.cpu(), torch.from_numpy(cpsd[f"it{it}_ball_slices"]))
    assert torch.equal(sel.j.cpu(), torch.from_numpy(cpsd[f"it{it}_ball_j"].astype(np.int64)))
    # per-point radii and a row subset
    rng = np.random.default_rng(it)
    idx = np.sort(rng.choice(len(cpsd["pos0"]), 700, replace=False))
    radii = (cpsd["d"] * rng.uniform(0.0, 2.0, len(idx))).astype(np.float32)
    radii[:3] = 0.0
    sub = p.selector.getPointsInRangeSelectionVectorized(torch.from_numpy(radii), cu(idx, torch.long))
    j, slices = O.ball_selection(cpsd["pos0"], cpsd[f"it{it}_pos_in"][idx], radii)
    assert np.array_equal(sub.slices.cpu().numpy(), slices) and np.array_equal(sub.j.cpu().numpy(), j)
    with pytest.raises(AssertionError):
        p.selector.getPointsInRangeSelectionVectorized(torch.ones(5), None)


def test_ball_selection_random_clouds(ng):
    rng = np.random.default_rng(5)
    tree = rng.normal(size=(20000, 3)).astype(np.float32)
    tree[:50] = tree[50:100]                                                   # duplicates
    query = np.concatenate([tree[:3000] + rng.normal(0, 0.01, (3000, 3)).astype(np.float32),
                            rng.uniform(-6, 6, (500, 3)).astype(np.float32)])  # some far outside the cloud: empty rows
    radii = rng.uniform(0.0, 0.35, len(query)).astype(np.float32)
    grid = ng._lib.Grid(cu(tree), 16)
    idx, off = grid.ball(cu(query), cu(radii))
    j, slices = O.ball_selection(tree, query, radii)
    assert np.array_equal(off.cpu().numpy(), slices) and np.array_equal(idx.cpu().numpy(), j)
    assert (np.diff(slices) == 0).any()


@pytest.mark.parametrize("it", [0, 1, 2])
def test_cpsd_decompositions_teacher_forced(ng, cpsd, it):
    t = f"it{it}_"
    p = _cpsd_processor(ng, cpsd, it)
    sel = p.selector.getPointsInRangeSelection(float(cpsd["d"]))
    nvt = p.decompositionor.getNormalFilteredNVT(sel, p.graph.n, 0.9)
    ref_w = cpsd[t + "nvt_eigval"]
    assert np.abs(nvt.eigval.cpu().numpy() - ref_w).max() <= 2e-6
    # smoothing from the reference's eigenpairs (signs are LAPACK's), then the point voting tensor on its smoothed normals
    dec_ref = ng.Decomposition(cu(cpsd[t + "nvt_eigval"]), cu(cpsd[t + "nvt_eigvec"]))
    f_n = dec_ref.getVUSmoothedNormals(p.graph.n)
    assert angle_between(f_n.cpu().numpy(), cpsd[t + "f_n"]).max() < 1e-4
    pvt = p.decompositionor.getNormalFilteredPVT(sel, cu(cpsd[t + "f_n"]), 0.9)
    ref2 = cpsd[t + "pvt_eigval"]
    assert np.abs(pvt.eigval.cpu().numpy() - ref2).max() <= 2e-6 * np.abs(ref2).max()
    lab = pvt.getVUFeatures(tau=0.3).cpu().numpy()
    # labels are thresholds on eigenvalues: only values within rounding of tau may differ
    differ = lab != cpsd[t + "classes"]
    assert differ.sum() <= 2 and (np.abs(ref2[differ] - 0.3).min(axis=1) < 1e-5).all()
    # our own eigenvectors: equal to LAPACK's up to sign where the spectrum is separated
    gap = np.minimum(ref2[:, 1] - ref2[:, 0], ref2[:, 2] - ref2[:, 1]) / np.abs(ref2).max()
    dots = np.abs((pvt.eigvec.cpu().numpy() * cpsd[t + "pvt_eigvec"]).sum(axis=1))
    assert (dots[gap > 1e-3] > 1 - 1e-4).all()


def test_cpsd_steps_on_ball_rows(ng, cpsd):
    p = _cpsd_processor(ng, cpsd, 0)
    n = len(cpsd["pos0"])
    sel = p.selector.getPointsInRangeSelection(float(cpsd["d"]))
    f_n, d = cu(cpsd["it0_f_n"]), float(cpsd["d"])
    edge = cu(np.ascontiguousarray(cpsd["it0_pvt_eigvec"][:, :, 0]))
    scale = np.abs(cpsd["pos0"]).max()
    for key, name in enumerate(("flat", "edge", "corner", "feature")):
        idx = (torch.arange(n, device="cuda") % 4 == key).nonzero().flatten()
        rows = sel.filter(idx)
        if name == "flat":
            new = p.denoiser.flat_step(rows, f_n, d, 0.5)
        elif name == "edge":
            new = p.denoiser.edge_step(rows, f_n, edge, d, 0.5)
        elif name == "corner":
            new = p.denoiser.corner_step(rows, f_n, d, 0.5)
        else:
            new = p.denoiser.feature_step(rows, f_n, d, 0.5)
        err = np.abs(new.cpu().numpy() - cpsd["ballrows_" + name]).max(axis=1) / scale
        # balls of 1-3 points give (near-)singular systems whose solution is rounding noise until the d-clamp cuts it
        assert (err > 1e-5).mean() < 0.01, (name, (err > 1e-5).mean(), err.max())


def test_cpsd_loop_vs_reference(ng, cpsd):
    """PostProcessing.ipynb#c9, method "CPSD": three iterations, free-running, through the mirror of the reference API."""
    p = _processor(ng, cpsd["pos0"], cpsd["n0"])
    g = p.graph
    d = float(cpsd["d"])
    l = ng.TorchUtils.averageEdgeLength(g.pos, p.selector.getKNNSelection(6).getEdgeIndex())
    assert abs(float(l) - float(cpsd["l"])) / float(cpsd["l"]) < 1e-6
    original_pos = g.pos.clone()
    alphas = [0.1, 1, 1]
    scale = np.abs(cpsd["pos0"]).max()
    for it in range(3):
        decomposition, f_n = p.getMartinFeatureDecomposition(r=d)
        classes = decomposition.getVUFeatures(tau=0.3)
        selection = p.selector.getKNNSelection(k=8)
        temp_pos = g.pos.clone()
        for key in range(3):
            indices = (classes == key).nonzero().flatten()
            if indices.size(0) == 0:
                continue
            elif key == 0:
                new_pos = p.denoiser.flat_step(selection.filter(indices), f_n, d * 20000, alphas[key])
            elif key == 1:
                new_pos = p.denoiser.edge_step(selection.filter(indices), f_n, decomposition.eigvec[..., 0], d * 20000, alphas[key])
            else:
                new_pos = p.denoiser.corner_step(selection.filter(indices), f_n, d * 20000, alphas[key])
            temp_pos[indices] = new_pos
        mask = (temp_pos - original_pos).norm(dim=1) < d
        g.pos[mask] = temp_pos[mask]
        g.n = f_n
        t = f"it{it}_"
        agree = (classes.cpu().numpy() == cpsd[t + "classes"]).mean()
        bad_n = (angle_between(f_n.cpu().numpy(), cpsd[t + "f_n"]) > 1e-4).mean()
        err = np.abs(g.pos.cpu().numpy() - cpsd[t + "pos_out"]).max(axis=1) / scale
        print(f"\nCPSD iteration {it}: labels agree {agree:.4%}, normals > 1e-4 rad {bad_n:.4%}, positions > 1e-5 {(err > 1e-5).mean():.4%}")
        # Free-running: eigenvector signs feed the smoothing.  Yardstick (tests/golden/noise_floor_cpsd.py, three trials):
        # the reference's own algorithm, re-run with its input normals moved by 1 ulp, agrees with itself on
        #   labels      99.89-99.94 % / 99.74-99.77 % / 98.55-98.92 %   after iterations 1 / 2 / 3,
        #   normals     0.37-0.42 %  / 1.9-2.4 %     / 6.7-8.8 %       further than 1e-4 rad,
        #   positions   1.7-1.9 %    / 6.3-8.0 %     / 14.3-17.2 %     further than 1e-5 of the extent.
        # The first iteration starts from identical inputs, so it must do better than that floor; the later ones must stay on it.
        floor_labels, floor_normals, floor_pos = (0.9985, 0.9965, 0.9840)[it], (0.0045, 0.03, 0.10)[it], (0.021, 0.09, 0.20)[it]   # (three trials: a little slack)
        assert agree > (0.999 if it == 0 else floor_labels), (it, agree)
        assert bad_n < (0.01 if it == 0 else floor_normals), (it, bad_n)
        assert (err > 1e-5).mean() < (0.02 if it == 0 else floor_pos), (it, (err > 1e-5).mean())


def test_tensor_division_is_correctly_rounded(ng):
    """The voting tensor divides six sums by the vote count with one shared reciprocal + remainder correction
    (point_math.cuh: CountDivider).  Every count 1..64 with random rows: bit-equal to IEEE division of the same sums."""
    rng = np.random.default_rng(11)
    n, reps = 4096, 64 * 200
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm.astype(np.float32)
    pos = rng.normal(size=(n, 3)).astype(np.float32)
    lens = np.tile(np.arange(1, 65), reps // 64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    j = rng.integers(0, n, off[-1]).astype(np.int32)
    rows = rng.integers(0, n, len(lens)).astype(np.int32)
    m = len(lens)
    ev = torch.empty((m, 3), device="cuda"); vec = torch.empty((m, 3, 3), device="cuda"); T = torch.empty((m, 3, 3), device="cuda")
    sw = torch.empty(m, dtype=torch.int32, device="cuda")
    lib = ng._lib.load()
    # x_thresh = 1: every neighbour votes (the quick test is off at the ends of the range, the exact sequence passes |x| <= 1)
    d_pos, d_nrm, d_j, d_off, d_rows = cu(pos), cu(nrm), cu(j), cu(off), cu(rows)          # keep the device copies alive
    ng._lib.check(lib.ngpd_nvt(d_pos.data_ptr(), d_nrm.data_ptr(), d_j.data_ptr(), d_off.data_ptr(), d_rows.data_ptr(), m, 0,
                               1.0, ev.data_ptr(), vec.data_ptr(), T.data_ptr(), sw.data_ptr(), None), "nvt")
    torch.cuda.synchronize()
    assert np.array_equal(sw.cpu().numpy(), lens)
    got = T.cpu().numpy()
    for L in range(1, 65):
        r = np.nonzero(lens == L)[0]
        nj = nrm[j[off[r][:, None] + np.arange(L)[None, :]]]                      # [rows, L, 3]
        outer = (nj[:, :, :, None] * nj[:, :, None, :]).astype(np.float32)
        acc = np.zeros((len(r), 3, 3), np.float32)
        for a in range(L):
            acc = (acc + outer[:, a]).astype(np.float32)
        want = (acc / np.float32(L)).astype(np.float32)
        assert np.array_equal(got[r], want), L


# ------------------------------------------------------------------------------------------------------
# mesh vertex update (Vertex_updating notebook)
# ------------------------------------------------------------------------------------------------------
def test_mesh_vertex_update(ng):
    m = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "mesh_update.npz")))
    scale = np.abs(m["v_noisy"]).max()
    for k in (1, 5):
        mesh = ng.Mesh(m["v_noisy"].copy(), m["f"])
        assert np.array_equal(mesh.getVertexTriangleAdjacency()[0], m["vta_faces"]) and np.array_equal(mesh.getVertexTriangleAdjacency()[1], m["vta_offsets"])
        mesh.updateVertices(m["face_normals"], k)
        assert np.abs(mesh.getVertices() - m[f"v_after_{k}"]).max() <= 1e-12 * scale, k
    # face normals as the reference computes them; the update pulls them towards the targets
    mesh = ng.Mesh(m["v_noisy"].copy(), m["f"])
    before = (mesh.getFaceNormals() * m["face_normals"]).sum(1).mean()
    mesh.updateVertices(m["face_normals"], 15)
    after = (mesh.getFaceNormals() * m["face_normals"]).sum(1).mean()
    assert after > before and after > 0.97
    # k = 0 leaves the vertices alone; a vertex without faces divides 0 by 0 like the reference (NaN), others are unaffected
    mesh0 = ng.Mesh(m["v_noisy"].copy(), m["f"]); mesh0.updateVertices(m["face_normals"], 0)
    assert np.array_equal(mesh0.getVertices(), m["v_noisy"])
    v_extra = np.concatenate([m["v_noisy"], [[0.0, 0.0, 0.0]]])
    mesh1 = ng.Mesh(v_extra, m["f"]); mesh1.updateVertices(m["face_normals"], 1)
    assert np.isnan(mesh1.getVertices()[-1]).all() and np.abs(mesh1.getVertices()[:-1] - m["v_after_1"]).max() <= 1e-12 * scale


# ------------------------------------------------------------------------------------------------------
# normal orientation on the GPU (GraphBuilder.flipNormals)
# ------------------------------------------------------------------------------------------------------
def _knn_graph(ng, pos, k):
    p = ng.Processor(ng.Pointcloud(cu(pos).clone()))
    p.graph.edge_index = p.graphBuilder.getKNNEdgeIndex(k)
    return p


def test_orientation_vs_reference(ng, fandisk):
    p = _knn_graph(ng, fandisk["pos0"], 12)
    assert np.array_equal(p.graph.edge_index[1].view(-1, 12).cpu().numpy(), fandisk["knn12_noself"])
    p.graph.n = cu(fandisk["n_pca"]).clone()
    p.graphBuilder.flipNormals()
    n_gpu = p.graph.n.cpu().numpy()
    assert np.array_equal(np.abs(n_gpu), np.abs(fandisk["n_pca"]))                    # only signs change
    agree = ((n_gpu * fandisk["n_flip"]).sum(1) > 0).mean()
    # the reference's unstable argsort picks one of many minimum trees (flat regions: thousands of edges of cost exactly 0);
    # the oracle's stable Kruskal agrees with it on 97-98 % of the signs, so does this tree
    print(f"\norientation agrees with the reference on {agree:.4%} of the points; {p.graphBuilder.orientation_info}")
    assert agree > 0.97
    assert p.graphBuilder.orientation_info["components"] == 1
    # the SciPy version of the same definition (another tie order)
    n_scipy = O.orient_normals_scipy(fandisk["pos0"], fandisk["n_pca"], p.graph.edge_index.cpu().numpy())
    assert ((n_gpu * n_scipy).sum(1) > 0).mean() > 0.97


def test_orientation_unique_tree_equals_oracle(ng):
    """Random normals make all edge costs distinct: the minimum spanning tree is unique and the propagated signs must be
    exactly those of the oracle's Kruskal + traversal."""
    rng = np.random.default_rng(8)
    pos = surface_cloud(8000, 8, noise=0.002)
    nrm = rng.normal(size=pos.shape); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm.astype(np.float32)
    p = _knn_graph(ng, pos, 10)
    nbr = p.graph.edge_index[1].view(-1, 10).cpu().numpy()
    p.graph.n = cu(nrm).clone()
    p.graphBuilder.flipNormals()
    want = O.orient_normals(pos, nrm, nbr)
    assert np.array_equal(p.graph.n.cpu().numpy(), want)
    assert p.graphBuilder.orientation_info["rounds"] <= 16


def test_orientation_properties(ng):
    """Two spheres far apart, normals radial with random signs: the sphere holding the top-most point comes out pointing
    outwards everywhere, the other one (not connected to the root) is left alone -- as the reference's DFS does."""
    rng = np.random.default_rng(9)
    u = rng.normal(size=(30000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    top, low = u[:20000], u[20000:] * 0.5 + np.array([0.0, 0.0, -10.0])
    pos = np.concatenate([top, low]).astype(np.float32)
    radial = np.concatenate([u[:20000], u[20000:]]).astype(np.float32)
    sign = rng.choice([-1.0, 1.0], len(pos)).astype(np.float32)
    nrm = radial * sign[:, None]
    p = _knn_graph(ng, pos, 8)
    p.graph.n = cu(nrm).clone()
    p.graphBuilder.flipNormals()
    out = p.graph.n.cpu().numpy()
    assert ((out[:20000] * radial[:20000]).sum(1) > 0).all()
    assert np.array_equal(out[20000:], nrm[20000:])
    assert p.graphBuilder.orientation_info["components"] >= 2
    # deterministic
    q = _knn_graph(ng, pos, 8)
    q.graph.n = cu(nrm).clone()
    q.graphBuilder.flipNormals()
    assert torch.equal(p.graph.n, q.graph.n)


# ------------------------------------------------------------------------------------------------------
# round 2: the notebook's "Ours" / "CTD-QEM" rows as fused session modes, digest, reservations, 10 M-point sampled check,
# sampler / noise on the device
# ------------------------------------------------------------------------------------------------------
def _ours_session(ng, ours, t, strategy, alphas, dmax, clamp):
    L = ng._lib
    sess = L.Session(cu(ours["pos0"]), 16)
    sess.set_state(cu(ours[t + "pos_in"]), cu(ours[t + "n_in"]))
    if clamp:
        sess.set_original(cu(ours["pos0"]))
    params = L.make_params(16, 8, None, 0.3, 3.0, 0.2, strategy, alphas, dmax, L.STEP_SNAPSHOT_CLASSES, clamp)
    sess.step(params)
    return sess.get_state(True)


@pytest.mark.parametrize("it", [0, 1])
def test_ours_clamp_mode_vs_reference(ng, ours, it):
    """PostProcessing.ipynb#c9 row "Ours" as ONE fused session step (NGPD_STEP_SNAPSHOT_CLASSES + clamp_radius), teacher-forced
    on the reference's inputs of each iteration: labels, smoothed normals, clamped positions."""
    L = ng._lib
    t = f"ours{it}_"
    d = float(ours["d"])
    pos, fn, lab = _ours_session(ng, ours, t, (L.STEP_FLAT, L.STEP_FEATURE, L.STEP_FEATURE), (1.0, 0.2, 1.0), d * 20000, d)
    lab, fn, pos = lab.cpu().numpy(), fn.cpu().numpy(), pos.cpu().numpy()
    same = lab == ours[t + "classes"]
    ang = angle_between(fn, ours[t + "f_n"])
    err = np.abs(pos - ours[t + "pos_out"]).max(axis=1) / np.abs(ours["pos0"]).max()
    print(f"\nOurs iteration {it}: labels differ {(~same).sum()}, normals > 1e-4 rad {(ang > 1e-4).mean():.4%}, positions > 1e-5 {(err > 1e-5).mean():.4%}")
    # yardstick of one iteration (test_session_labels_vs_reference): the reference against itself under a 1-ulp change
    assert (~same).sum() <= 7 and (ang > 1e-4).mean() < 0.0080 and (err > 1e-5).mean() < 0.0377
    # rows whose stage-1 normals agree with the reference's in their whole update neighbourhood must agree to 1e-5
    good = ang <= 1e-4
    nbr = ng._lib.Grid(cu(ours["pos0"]), 8).knn(cu(ours[t + "pos_in"]), 8).cpu().numpy()
    clean = good[nbr].all(axis=1) & same
    assert clean.mean() > 0.95 and err[clean].max() < 1e-5, (clean.mean(), err[clean].max())


def test_ctd_qem_mode_vs_reference(ng, ours):
    """row "CTD-QEM": feature_step on every point from one snapshot = strategy feature/feature/feature in snapshot mode"""
    L = ng._lib
    pos, fn, _ = _ours_session(ng, ours, "qem0_", (L.STEP_FEATURE,) * 3, (1.0, 1.0, 1.0), float(ours["d"]), 0.0)
    ang = angle_between(fn.cpu().numpy(), ours["qem0_f_n"])
    err = np.abs(pos.cpu().numpy() - ours["qem0_pos_out"]).max(axis=1) / np.abs(ours["pos0"]).max()
    assert (ang > 1e-4).mean() < 0.0080 and (err > 1e-5).mean() < 0.0377
    nbr = L.Grid(cu(ours["pos0"]), 8).knn(cu(ours["qem0_pos_in"]), 8).cpu().numpy()
    clean = (ang <= 1e-4)[nbr].all(axis=1)
    assert err[clean].max() < 1e-5


def test_snapshot_mode_equals_operator_composition(ng, ours):
    """the fused "Ours" step against the same loop composed from the public operators (as the notebook writes it)"""
    L = ng._lib
    p = _processor(ng, ours["ours0_pos_in"], ours["ours0_n_in"], tree=ours["pos0"])
    g = p.graph
    d = float(ours["d"])
    original = cu(ours["pos0"])
    decomposition, f_n = p.getMyFeatureDecomposition()
    classes = decomposition.getClasses()
    selection = p.selector.getKNNSelection(8)
    temp_pos = g.pos.clone()
    for key in range(3):
        indices = (classes == key).nonzero().flatten()
        if indices.size(0) == 0:
            continue
        if key == 0:
            new_pos = p.denoiser.flat_step(selection.filter(indices), f_n, d * 20000, [1, 0.2, 1][key])
        else:
            new_pos = p.denoiser.feature_step(selection.filter(indices), f_n, d * 20000, [1, 0.2, 1][key])
        temp_pos[indices] = new_pos
    mask = (temp_pos - original).norm(dim=1) < d
    g.pos[mask] = temp_pos[mask]
    pos, fn, lab = _ours_session(ng, ours, "ours0_", (L.STEP_FLAT, L.STEP_FEATURE, L.STEP_FEATURE), (1.0, 0.2, 1.0), d * 20000, d)
    assert torch.equal(lab.long(), classes.long())
    assert float((pos - g.pos).abs().max()) / float(original.abs().max()) < 2e-6
    assert angle_between(fn.cpu().numpy(), f_n.cpu().numpy()).max() < 1e-4


def test_checksum_is_order_independent_and_reserve_changes_nothing(ng):
    L = ng._lib
    n = 60000
    pts = cu(surface_cloud(n, seed=3))
    nrm = torch.nn.functional.normalize(torch.randn((n, 3), generator=torch.Generator().manual_seed(1)), dim=1).cuda()
    a = L.Session(pts, 16)
    a.set_state(pts, nrm)
    s, c = a.mean_edge_length_parts(6)
    params = L.make_params(dmax=2.0 * s / c)
    for _ in range(2):
        a.step(params)
    da = a.checksum()
    # the same cloud handed over in another order, with the permutation as global ids, and everything reserved up front
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(2)).cuda()
    b = L.Session(pts[perm].contiguous(), 16)
    b.reserve(16)
    b.set_state(pts[perm].contiguous(), nrm[perm].contiguous())
    for _ in range(2):
        b.step(params)
    db = b.checksum(perm)
    assert da == db, (da, db)
    assert da[10] == n and da[6] + da[7] + da[8] == n and da[9] == 0
    pa, na, la = a.get_state(True)
    assert da[6] == int((la == 0).sum()) and da[7] == int((la == 1).sum())
    assert abs(da[2] - (1 << 64) * (da[2] >= (1 << 63)) - float(pa[:, 0].double().sum()) * 2 ** 24) < n
    # a single changed bit changes the digest
    pa2 = pa.clone(); pa2[17, 1] = torch.nextafter(pa2[17, 1], torch.tensor(9.0, device="cuda"))
    a.set_state(pa2, na)
    assert a.checksum()[0] != da[0]


def test_step_at_10M_points_sampled_against_oracle(ng):
    """A full-size step checked where the oracle can follow: sampled rows of a 10 M-point run against the exact search, an fp64
    brute force and the NumPy oracle (bench.py's `validated` leg)."""
    import argparse
    import bench
    L = ng._lib
    dev = torch.device("cuda", 0)
    args = argparse.Namespace(surface="creased", strategy="flat/edge/feature", clamp=False)
    n = 10_000_000
    noisy, analytic, _ = bench.make_shard(args, n, dev, 0, 1)
    nrm = bench.single_gpu_normals(noisy, analytic)
    sess = L.Session(noisy, 16)
    sess.set_state(noisy, nrm)
    s, c = sess.mean_edge_length_parts(6)
    params = L.make_params(dmax=2.0 * s / c)
    for _ in range(3):
        sess.step(params)                                       # iteration 3 answers most rows from stored candidates (tier 0)
    v = bench.validate(args, sess, params, noisy, n, dev, 4000, 48)
    print("\n10 M-point step:", v)
    assert v["knn_rows_equal_exact_search"] == v["rows_sampled"] == 4000
    assert v["knn_rows_equal_bruteforce"] == v["knn_rows_bruteforce_fp64"] == 48
    assert v["labels_equal_oracle"] >= 3996
    assert v["smoothed_normals_within_1e-4_rad_of_oracle"] >= 3960


def test_sampler_and_noise_on_device(ng, tmp_path):
    """Pointcloud.sampleObj / Noise.generateNoise with CUDA tensors: samples lie on the mesh in proportion to
    the face areas, are reproducible under the global seed, and the noise has the requested level along the normals."""
    obj = tmp_path / "two_squares.obj"
    obj.write_text("\n".join(["v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 3 0 0", "v 3 2 0", "v 3 2 2", "v 3 0 2",
                               "f 1 2 3", "f 1 3 4", "f 5 6 7", "f 5 7 8"]) + "\n")
    torch.manual_seed(0); torch.cuda.manual_seed(0)
    pc = ng.Pointcloud.sampleObj(str(obj), 2_000_000, device="cuda")
    assert pc.v.is_cuda and pc.n.is_cuda and pc.v.shape == (2_000_000, 3)
    v, nr = pc.v, pc.n
    small, big = v[:, 2].abs() < 1e-6, (v[:, 0] - 3).abs() < 1e-6
    assert bool((small | big).all())
    assert abs(float(big.float().mean()) - 0.8) < 0.002
    assert float((v[small, :2].mean(0) - 0.5).abs().max()) < 0.002 and float((v[big, 1:].mean(0) - 1.0).abs().max()) < 0.004
    assert bool((nr[small].abs() - torch.tensor([0.0, 0.0, 1.0], device="cuda")).abs().max() < 1e-6)
    torch.manual_seed(0); torch.cuda.manual_seed(0)
    again = ng.Pointcloud.sampleObj(str(obj), 2_000_000, device="cuda")
    assert torch.equal(again.v, pc.v)
    # noise along the normal on a CUDA graph: sigma = level x mean edge length, positions move only along n
    p = ng.Processor(ng.Pointcloud(pc.v.clone(), pc.n.clone()))
    torch.manual_seed(5)
    p.noise.generateNoise(0.3, 0.01, keepNormals=True)
    off = p.graph.pos - p.graph.gt
    along = (off * pc.n).sum(1)
    assert float((off - along[:, None] * pc.n).abs().max()) < 1e-6
    assert abs(float(along.std()) - 0.003) < 3e-5 and abs(float(along.mean())) < 2e-5


def test_pruned_delta_pass_is_exact(ng):
    """flat_step's delta (Denoiser.py:107) from the per-block maxima of the stage-2 pass + a look at the candidate blocks only
    (session_class_max_pruned_kernel) must be the exact maximum over the class-0 neighbour multiset, in every iteration -- the
    reference centre it prunes with is the previous iteration's (the bounding-box centre the first time)."""
    import ctypes
    L = ng._lib
    lib = L.load()
    n = 300000
    pts = cu(surface_cloud(n, seed=5, noise=0.003))
    # off-centre cloud: the bounding-box centre is far from the centroid, so the first iteration prunes with a poor reference
    pts = torch.cat([pts, pts[:20000] * 0.1 + torch.tensor([4.0, 0.5, -0.25], device="cuda")]).contiguous()
    n = pts.size(0)
    nrm = torch.nn.functional.normalize(torch.randn((n, 3), generator=torch.Generator().manual_seed(3)), dim=1).cuda()
    sess = L.Session(pts, 16)
    sess.set_state(pts, nrm)
    s, c = sess.mean_edge_length_parts(6)
    params = L.make_params(dmax=2.0 * s / c)
    ref = ctypes.byref(params)
    view = lambda which, shape, ts: torch.as_tensor(type("V", (), {"__cuda_array_interface__": {"data": (int(lib.ngpd_session_buffer(sess._h, which)), False),
                                                                                           "shape": shape, "typestr": ts, "version": 2}})(), device="cuda")
    perm = sess.order().long()
    for it in range(4):
        L.check(lib.ngpd_session_phase_features(sess._h, ref, 0, L.stream()), "f0")
        L.check(lib.ngpd_session_phase_features(sess._h, ref, 1, L.stream()), "f1")
        L.check(lib.ngpd_session_phase_flat_scalars(sess._h, ref, 0, 0, L.stream()), "s0")
        L.check(lib.ngpd_session_phase_flat_scalars(sess._h, ref, 0, 1, L.stream()), "s1")
        cd = view(4, (4,), "<f4").clone()
        lab = view(5, (n,), "|u1")
        idx = view(6, (n, 16), "<i4")
        pos4 = view(0, (n, 4), "<f4")
        rows = (lab == 0).nonzero().flatten()
        nb = idx[rows][:, :8].long().reshape(-1)
        d = pos4[nb, :3] - cd[:3]
        # torch's norm of 3 components on the GPU is not the fma chain the kernels (and torch-CPU) use: compare with a relative bound
        exact = float(d.double().norm(dim=1).max())
        centre = pos4[nb, :3].double().mean(0)
        assert float((centre - cd[:3].double()).abs().max()) < 1e-6 * float(pts.abs().max())
        assert abs(float(cd[3]) - exact) <= 2e-7 * exact, (it, float(cd[3]), exact)
        for key in range(3):
            L.check(lib.ngpd_session_phase_update(sess._h, ref, key, L.stream()), "u")
        L.check(lib.ngpd_session_phase_commit_normals(sess._h), "c")
