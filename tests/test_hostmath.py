"""The product's per-point math headers (csrc/point_math.cuh, csrc/eig3.cuh) compiled for the host with g++ and
checked against the reference's golden vectors: lets the kernels' arithmetic be verified without a GPU.  The
harness (tests/hostmath/) is test-only; the CUDA kernels include the same headers."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

import ngpd_oracle as O
from conftest import ROOT, angle_between

HM = os.path.join(ROOT, "tests", "hostmath")


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HM, "libhostmath.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", os.path.join(HM, "hostmath.cpp"), "-o", so],
                   check=True)
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


XT = None


def xt():
    global XT
    if XT is None:
        XT = ctypes.c_float(float(O.acos_threshold(math.pi * 5 / 12)))
    return XT


@pytest.mark.parametrize("it", [0, 1])
def test_tensor_eig_smooth_labels(hm, fandisk, it):
    t = f"it{it}_"
    n = len(fandisk["pos0"])
    pos, nrm, idx = (np.ascontiguousarray(fandisk[t + k]) for k in ("pos_in", "n_in", "knn16"))
    w = np.zeros((n, 3), np.float32); V = np.zeros((n, 3, 3), np.float32); T = np.zeros((n, 3, 3), np.float32); sw = np.zeros(n, np.int32)
    hm.hm_nvt(P(pos), P(nrm), P(idx), None, ctypes.c_int64(n), 16, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, fandisk[t + "T1"])                                   # voting tensors bit-exact
    assert np.abs(w - fandisk[t + "eigval1"]).max() < 1e-6
    gap = np.minimum(np.diff(fandisk[t + "eigval1"], axis=1).min(axis=1), 1.0)
    signs = ((V * fandisk[t + "eigvec1"]).sum(1) > 0).all(1)
    assert signs[gap > 1e-2].all()                                                # LAPACK sign convention reproduced
    assert signs.mean() > 0.99
    good = signs & (gap > 1e-2)
    assert np.abs(V - fandisk[t + "eigvec1"])[good].max() < 1e-4
    # smoothing fed the reference's eigenvectors: bit-exact
    out = np.zeros((n, 3), np.float32)
    gw, gV = np.ascontiguousarray(fandisk[t + "eigval1"]), np.ascontiguousarray(fandisk[t + "eigvec1"])
    hm.hm_smooth(P(gw), P(gV), P(nrm), ctypes.c_int64(n), ctypes.c_float(0.3), ctypes.c_float(3.0), P(out))
    assert np.array_equal(out, fandisk[t + "f_n"])
    # free running (own eigenvectors): within the reference's own 1-ulp noise floor (0.82 % > 1e-4 rad)
    hm.hm_smooth(P(w), P(V), P(nrm), ctypes.c_int64(n), ctypes.c_float(0.3), ctypes.c_float(3.0), P(out))
    assert (angle_between(out, fandisk[t + "f_n"]) > 1e-4).mean() < 0.0082
    # stage 2 + labels
    fn = np.ascontiguousarray(fandisk[t + "f_n"])
    hm.hm_nvt(P(pos), P(fn), P(idx), None, ctypes.c_int64(n), 16, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, fandisk[t + "T2"])
    lab = np.zeros(n, np.uint8)
    hm.hm_classify(P(w), ctypes.c_int64(n), ctypes.c_float(0.2), P(lab))
    assert np.array_equal(lab, fandisk[t + "classes"])                            # labels bit-exact from own eigenvalues


def test_tensor_eig_labels_k32(hm, fandisk, fandisk_k32):
    """The kernels' per-point math at BASELINE configs[2]'s neighbourhood size (k = 32), against the reference's tensors."""
    g = fandisk_k32
    n = len(fandisk["pos0"])
    pos, nrm, idx = (np.ascontiguousarray(a) for a in (fandisk["pos0"], fandisk["n_flip"], g["knn32"]))
    w = np.zeros((n, 3), np.float32); V = np.zeros((n, 3, 3), np.float32); T = np.zeros((n, 3, 3), np.float32); sw = np.zeros(n, np.int32)
    hm.hm_nvt(P(pos), P(nrm), P(idx), None, ctypes.c_int64(n), 32, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, g["T1"])
    assert np.abs(w - g["eigval1"]).max() < 1e-6
    out = np.zeros((n, 3), np.float32)
    hm.hm_smooth(P(w), P(V), P(nrm), ctypes.c_int64(n), ctypes.c_float(0.3), ctypes.c_float(3.0), P(out))
    assert (angle_between(out, g["f_n"]) > 1e-4).mean() < 0.0082              # the reference's own 1-ulp noise floor
    fn = np.ascontiguousarray(g["f_n"])
    hm.hm_nvt(P(pos), P(fn), P(idx), None, ctypes.c_int64(n), 32, xt(), P(w), P(V), P(T), P(sw))
    assert np.array_equal(T, g["T2"])
    lab = np.zeros(n, np.uint8)
    hm.hm_classify(P(w), ctypes.c_int64(n), ctypes.c_float(0.2), P(lab))
    assert np.array_equal(lab, g["classes"])


@pytest.mark.parametrize("it", [0, 1])
def test_updates(hm, fandisk, it):
    t = f"it{it}_"
    cur = np.ascontiguousarray(fandisk[t + "pos_in"]).copy()
    fn = np.ascontiguousarray(fandisk[t + "f_n"]); cls = fandisk[t + "classes"]; idx8 = fandisk[t + "knn8"]
    edge = np.ascontiguousarray(fandisk[t + "eigvec2"][:, :, 0])
    d = ctypes.c_float(float(np.float32(2) * fandisk["l"]))
    for key, kind, alpha in ((0, 0, 1.0), (1, 1, 0.2), (2, 2, 1.0)):
        rows = np.nonzero(cls == key)[0].astype(np.int32)
        sub = np.ascontiguousarray(idx8[rows])
        _, delta = O.flat_center_delta(cur, sub)
        out = np.zeros((len(rows), 3), np.float32)
        hm.hm_update(kind, P(cur), P(fn), P(edge), P(sub), P(rows), ctypes.c_int64(len(rows)), 8, ctypes.c_float(alpha), d,
                     ctypes.c_float(float(delta)), P(out))
        ref = fandisk[t + f"pos_after_class{key}"]
        assert np.abs(out - ref[rows]).max() / np.abs(ref).max() < 1e-5
        cur = ref.copy()


def test_pca_normals(hm, fandisk):
    n = len(fandisk["pos0"])
    pos, idx = np.ascontiguousarray(fandisk["pos0"]), np.ascontiguousarray(fandisk["knn12_noself"])
    out = np.zeros((n, 3), np.float32)
    hm.hm_pca(P(pos), P(idx), ctypes.c_int64(n), 12, P(out))
    assert angle_between(out, fandisk["n_pca"]).max() < 1e-4                      # sign-consistent with LAPACK's


def test_eigh3_properties(hm):
    """Identities from the reference's PatchGeneration/Tests/test_RotationMatrix.py:86-118 on our solver:
    (n (x) n) has top eigenvector +-n with eigenvalue |n|^2, eigenvectors are orthonormal, A V = V diag(w)."""
    rng = np.random.default_rng(3)
    m = 2000
    v = rng.normal(size=(m, 3)).astype(np.float32)
    T = (v[:, :, None] * v[:, None, :]).astype(np.float32)
    w = np.zeros((m, 3), np.float32); V = np.zeros((m, 3, 3), np.float32)
    hm.hm_eigh3(P(np.ascontiguousarray(T)), ctypes.c_int64(m), P(w), P(V))
    top = V[:, :, 2]
    un = v / np.linalg.norm(v, axis=1, keepdims=True)
    assert np.abs(np.abs((top * un).sum(1)) - 1).max() < 1e-5
    assert np.abs(w[:, 2] - (v * v).sum(1)).max() / (v * v).sum(1).max() < 1e-5
    A = rng.normal(size=(m, 3, 3)).astype(np.float32); A = ((A + A.transpose(0, 2, 1)) / 2).astype(np.float32)
    hm.hm_eigh3(P(np.ascontiguousarray(A)), ctypes.c_int64(m), P(w), P(V))
    assert (np.diff(w, axis=1) >= 0).all()
    I = np.einsum("nij,nik->njk", V, V)
    assert np.abs(I - np.eye(3)).max() < 1e-5
    assert np.abs(np.einsum("nij,njk->nik", A, V) - V * w[:, None, :]).max() < 2e-5
    wr = np.linalg.eigvalsh(A.astype(np.float64))
    assert np.abs(w - wr).max() < 1e-5


def test_eigh3_static_equals_generic_transcription(hm):
    """The product's n = 3 specialisation (csrc/eig3.cuh: named registers, one sweep body for QL and QR) is
    bit-identical to the array-indexed transcription of the LAPACK loops (tests/hostmath/eig3_generic.h) on
    general, voting-tensor-like, rank-deficient, diagonal, structurally sparse, scaled and repeated-eigenvalue inputs."""
    rng = np.random.default_rng(0)
    m = 200_000

    def sym(a):
        return (a + a.transpose(0, 2, 1)) / 2

    def votes(k, spread):
        base = rng.normal(size=(m, 1, 3)); base /= np.linalg.norm(base, axis=2, keepdims=True)
        n = base + spread * rng.normal(size=(m, k, 3)); n /= np.linalg.norm(n, axis=2, keepdims=True)
        n = n.astype(np.float32)
        return np.einsum("mki,mkj->mij", n, n) / np.float32(k)

    cases = {"general": sym(rng.normal(size=(m, 3, 3))), "votes_flat": votes(16, 0.05), "votes_noisy": votes(16, 0.5),
             "votes_rank1": votes(1, 0.0), "votes_rank2": votes(2, 0.3), "zero": np.zeros((4, 3, 3)), "identity": np.tile(np.eye(3), (4, 1, 1)),
             "tiny": sym(rng.normal(size=(m, 3, 3))) * 1e-20, "huge": sym(rng.normal(size=(m, 3, 3))) * 1e15}
    ax = np.eye(3)[rng.integers(0, 3, (m, 16))] * rng.choice([-1, 1], (m, 16, 1))
    cases["axis_aligned"] = np.einsum("mki,mkj->mij", ax, ax) / 16
    d = np.zeros((m, 3, 3)); d[:, 0, 0], d[:, 1, 1], d[:, 2, 2] = rng.normal(size=(3, m)); cases["diagonal"] = d
    for name, (r, c) in {"a31_zero": (2, 0), "a21_zero": (1, 0), "a32_zero": (2, 1)}.items():
        a = sym(rng.normal(size=(m, 3, 3))); a[:, r, c] = a[:, c, r] = 0; cases[name] = a
    pts = rng.normal(size=(m, 12, 3)) * np.array([1, 1, 0.01]); pts -= pts.mean(1, keepdims=True)
    cases["covariance"] = np.einsum("mki,mkj->mij", pts, pts)
    q = np.linalg.qr(rng.normal(size=(m, 3, 3)))[0]
    lam = np.stack([np.ones(m), np.ones(m), rng.uniform(0, 2, m)], 1)
    cases["repeated"] = np.einsum("mij,mj,mkj->mik", q, lam, q)
    for name, T in cases.items():
        T = np.ascontiguousarray(T.astype(np.float32)); n = len(T)
        w1 = np.zeros((n, 3), np.float32); V1 = np.zeros((n, 3, 3), np.float32); w2 = w1.copy(); V2 = V1.copy()
        hm.hm_eigh3(P(T), ctypes.c_int64(n), P(w1), P(V1))
        hm.hm_eigh3_generic(P(T), ctypes.c_int64(n), P(w2), P(V2))
        assert np.array_equal(w1.view(np.uint32), w2.view(np.uint32)), name
        assert np.array_equal(V1.view(np.uint32), V2.view(np.uint32)), name


@pytest.mark.parametrize("thresh", [0.25881898403167725, 0.5, 0.9999, 1.0, 0.0, -1.0, 1e-4])
def test_nvt_weight_shortcut_equals_exact_sequence(hm, thresh):
    """The division-free neighbour filter gives the reference's 0/1 weight on random edges, on edges constructed to sit
    within a few ulp of the threshold, on zero-length and tiny edges, and for degenerate thresholds."""
    rng = np.random.default_rng(7)
    m = 400_000
    vi = rng.normal(size=(m, 3)).astype(np.float32)
    dv = (rng.normal(size=(m, 3)) * 10.0 ** rng.uniform(-6, 1, (m, 1))).astype(np.float32)
    nj = rng.normal(size=(m, 3)); nj /= np.linalg.norm(nj, axis=1, keepdims=True)
    # second half: normals with |u.n| = thresh * (1 + tiny), i.e. decisions on the edge
    h = m // 2
    u = dv[h:].astype(np.float64); u /= np.maximum(np.linalg.norm(u, axis=1, keepdims=True), 1e-300)
    t = rng.normal(size=(m - h, 3)); t -= (t * u).sum(1, keepdims=True) * u; t /= np.linalg.norm(t, axis=1, keepdims=True)
    x = np.clip(abs(thresh) * (1 + rng.uniform(-3e-6, 3e-6, (m - h, 1))), 0, 1)
    nj[h:] = x * u * rng.choice([-1, 1], (m - h, 1)) + np.sqrt(1 - x * x) * t
    nj = nj.astype(np.float32)
    dv[:50] = 0; dv[50:100] *= 1e-20
    vj = (vi + dv).astype(np.float32)
    ex = np.zeros(m, np.uint8); qk = np.zeros(m, np.uint8)
    hm.hm_weight(P(vi), P(vj), P(nj), ctypes.c_int64(m), ctypes.c_float(thresh), P(ex), P(qk))
    assert np.array_equal(ex, qk)
    if 0 < thresh < 1:
        assert 0.05 < ex[h:].mean() < 0.95          # the constructed edges really straddle the threshold


def test_sorting_networks_zero_one_principle(hm):
    """csrc/sortnet.cuh: a comparator network sorts every input iff it sorts every 0/1 input.  All 2^16 inputs of the
    16-key sorter and all 2^8 of the 8-key one; the 32-key sorter and the bitonic mergers on random and adversarial keys."""
    rng = np.random.default_rng(0)

    def run(keys, bitonic=False):
        a = np.ascontiguousarray(keys, dtype=np.uint32)
        assert hm.hm_sortnet(P(a), len(a), int(bitonic)) == 0
        return a

    for n in (8, 16):
        for bits in range(1 << n):
            v = np.array([(bits >> i) & 1 for i in range(n)], dtype=np.uint32)
            out = run(v)
            assert np.array_equal(out, np.sort(v)), (n, bits)
    for _ in range(2000):
        v = rng.integers(0, 2 ** 32, 32, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(run(v), np.sort(v))
        v01 = rng.integers(0, 2, 32).astype(np.uint32)
        assert np.array_equal(run(v01), np.sort(v01))
    # bitonic merge: ascending run followed by a descending run of any split
    for n in (8, 16, 32):
        for split in range(n + 1):
            for _ in range(50):
                v = rng.integers(0, 1000, n).astype(np.uint32)
                v = np.concatenate([np.sort(v[:split]), np.sort(v[split:])[::-1]])
                assert np.array_equal(run(v, True), np.sort(v)), (n, split)


def test_cpsd_normal_filtered_tensors(hm, cpsd):
    """csrc/point_math.cuh nvt_normal_point / pvt_normal_point (Yadav-2018 baseline) on the reference's radius selection:
    tensors and eigenvalues against the reference's own (recorded) ones."""
    n = len(cpsd["pos0"])
    x_le = ctypes.c_float(float(O.acos_threshold_le(0.9)))
    for it in range(2):
        t = f"it{it}_"
        j = np.ascontiguousarray(cpsd[t + "ball_j"], dtype=np.int32)
        off = np.ascontiguousarray(cpsd[t + "ball_slices"], dtype=np.int32)
        for mode, nrm_key, T_key, w_key, V_key in ((0, "n_in", "T_nvt", "nvt_eigval", "nvt_eigvec"), (1, "f_n", "T_pvt", "pvt_eigval", "pvt_eigvec")):
            pos = np.ascontiguousarray(cpsd[t + "pos_in"], dtype=np.float32)
            nrm = np.ascontiguousarray(cpsd[t + nrm_key], dtype=np.float32)
            w = np.empty((n, 3), np.float32); V = np.empty((n, 3, 3), np.float32); T = np.empty((n, 3, 3), np.float32)
            sw = np.empty(n, np.int32)
            hm.hm_normal_filtered(mode, P(pos), P(nrm), P(j), P(off), ctypes.c_int64(n), x_le, P(w), P(V), P(T), P(sw))
            ref_T, ref_w = cpsd[t + T_key], cpsd[t + w_key]
            if mode == 0:
                assert np.array_equal(T, ref_T)
            else:
                assert np.abs(T - ref_T).max() <= 2e-6 * np.abs(ref_T).max()
            assert np.abs(w - ref_w).max() <= 2e-6 * np.abs(ref_w).max()
            # eigenvectors up to sign where the spectrum is well separated
            gap = np.minimum(ref_w[:, 1] - ref_w[:, 0], ref_w[:, 2] - ref_w[:, 1]) / np.abs(ref_w).max()
            ok = gap > 1e-3
            dots = np.abs((V * cpsd[t + V_key]).sum(axis=1))
            assert (dots[ok] > 1 - 1e-4).all()


def _random_voting_tensors(n, k, rng, spread):
    """T = sum w n n^T / sum w of k unit normals scattered around up to three directions (flat / crease / corner mixtures)"""
    base = rng.normal(size=(n, 3, 3))
    base /= np.linalg.norm(base, axis=2, keepdims=True)
    which = rng.integers(0, 3, (n, k)) % rng.integers(1, 4, (n, 1))
    nrm = base[np.arange(n)[:, None], which] + rng.normal(scale=spread, size=(n, k, 3))
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    w = (rng.random((n, k)) < 0.7).astype(np.float32)
    w[w.sum(1) == 0] = 1
    nrm = nrm.astype(np.float32)
    T = np.einsum("nk,nki,nkj->nij", w, nrm, nrm).astype(np.float32) / w.sum(1)[:, None, None]
    T = ((T + T.transpose(0, 2, 1)) * np.float32(0.5)).astype(np.float32)
    return np.ascontiguousarray(T)


def _fast_vs_lapack(hm, T):
    n = len(T)
    lab = np.zeros(n, np.uint8); cert = np.zeros(n, np.uint8); l3 = np.zeros(n, np.float32); vec = None
    hm.hm_classify_fast(P(T), ctypes.c_int64(n), ctypes.c_float(0.2), P(lab), P(cert), P(l3))
    w = np.zeros((n, 3), np.float32); V = np.zeros((n, 3, 3), np.float32)
    hm.hm_eigh3(P(T), ctypes.c_int64(n), P(w), P(V))
    ref = np.zeros(n, np.uint8)
    hm.hm_classify(P(w), ctypes.c_int64(n), ctypes.c_float(0.2), P(ref))
    return lab, cert.astype(bool), l3, vec, w, V, ref


def test_fast_labels_agree_with_lapack_order(hm, fandisk, fandisk_k32):
    """csrc/eig3_fast.cuh: wherever the closed-form path calls its label certain it equals the label from the LAPACK-order
    eigenvalues (the rest is recomputed by that path in the kernel) and it is certain almost everywhere."""
    rng = np.random.default_rng(7)
    sets = {"fandisk it0": fandisk["it0_T2"], "fandisk it1": fandisk["it1_T2"], "fandisk k32": fandisk_k32["T2"] if "T2" in fandisk_k32 else fandisk["it0_T1"],
            "random tight": _random_voting_tensors(400000, 16, rng, 0.02), "random loose": _random_voting_tensors(400000, 16, rng, 0.3),
            "random k32": _random_voting_tensors(200000, 32, rng, 0.1)}
    for name, T in sets.items():
        T = np.ascontiguousarray(T.astype(np.float32))
        lab, cert, l3, vec, w, V, ref = _fast_vs_lapack(hm, T)
        wrong = cert & (lab != ref)
        print(f"{name}: {len(T)} tensors, uncertain {1 - cert.mean():.3e}, labels {np.bincount(ref, minlength=3).tolist()}, "
              f"max |l3 - lapack| {np.abs(l3 - w[:, 0])[cert].max():.2e}")
        assert wrong.sum() == 0, (name, int(wrong.sum()))
        assert cert.mean() > 0.995, (name, cert.mean())
        assert np.abs(l3 - w[:, 0])[cert].max() < 3e-6
    # rows pushed onto the decision boundaries: uncertain is allowed, a certain wrong answer is not
    T = _random_voting_tensors(300000, 16, rng, 0.15)
    lab, cert, l3, vec, w, V, ref = _fast_vs_lapack(hm, T)
    l1, l2, l3_ = w[:, 2], w[:, 1], w[:, 0]
    g = np.stack([0.2 * (l1 - l2), l2 - l3_, l3_], 1)
    srt = np.sort(g, axis=1)
    close = (srt[:, 2] - srt[:, 1]) < 1e-4
    assert (cert & (lab != ref)).sum() == 0
    print(f"near-boundary rows: {close.sum()}, of them certain {cert[close].mean():.3f}")
    # degenerate input goes to the fallback instead of producing a label
    bad = np.zeros((4, 3, 3), np.float32)
    bad[1] = np.eye(3, dtype=np.float32) / 3
    bad[2, 0, 0] = np.nan
    bad[3] = np.float32(1e30)
    _, cert, *_ = _fast_vs_lapack(hm, bad)
    assert not cert.any()
