"""CPU oracle for the normal-guided point-cloud denoising hot path.

TEST INFRASTRUCTURE ONLY.  This module is the checker the CUDA path is compared against; it is
imported by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl
reference` legs, and by nothing in the product package.

It restates, in NumPy (fp32 arrays, explicit per-edge formulas), the algorithm of the reference's
`Pointcloud/Modules` hot path.  Every function cites the reference file:line it follows.  Two
third-party pieces of arithmetic are not in the reference checkout and are taken from the same
libraries the reference calls:
  * `scipy.spatial.cKDTree.query`  (reference: Selector.py:141,243; pinned scipy=1.13.1 in
    environment_optimal.yml:178; here 1.18.x) -- k nearest tree points by fp64 squared distance of the
    fp32-upcast coordinates.  `knn_bruteforce` below restates its published semantics exactly with the
    tie rule (distance, index) that SciPy leaves unspecified.
  * LAPACK `ssyevd` through `torch.linalg.eigh` (reference: Decompositionor.py:300, GraphBuilder.py:110;
    pinned pytorch=1.12.1 + mkl=2024.1.0; here torch 2.11 + MKL 2024.2).  Eigenvector signs are whatever
    that LAPACK returns; the reference's smoothing step depends on them.

Pinned: `tests/test_oracle_golden.py` checks every function here against golden vectors recorded from
the unmodified reference (imported in the build container with stub modules for its missing
dependencies; generator committed as tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# kNN (Selector.getKNNSelection, Selector.py:235-246; GraphBuilder.getKNNEdgeIndex, GraphBuilder.py:60-63)
# ------------------------------------------------------------------------------------------------
def knn_bruteforce(tree: np.ndarray, query: np.ndarray, k: int, chunk: int = 2048) -> np.ndarray:
    """k nearest rows of `tree` for every row of `query`: fp64 ((dx^2+dy^2)+dz^2) of the fp32-upcast
    coordinates, ascending, ties broken by the lower tree index.  Returns int64 [m,k]; slots beyond the
    tree size hold len(tree) (what SciPy returns for missing neighbours)."""
    t = np.asarray(tree, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    n, m = len(t), len(q)
    out = np.full((m, k), n, dtype=np.int64)
    kk = min(k, n)
    for s in range(0, m, chunk):
        qs = q[s:s + chunk]
        dx = qs[:, None, 0] - t[None, :, 0]
        dy = qs[:, None, 1] - t[None, :, 1]
        dz = qs[:, None, 2] - t[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        # lexicographic (d2, index): stable argsort on d2 keeps index order among equals
        order = np.argsort(d2, axis=1, kind="stable")[:, :kk]
        out[s:s + chunk, :kk] = order
    return out


def knn_kdtree(tree: np.ndarray, query: np.ndarray, k: int, workers: int = 1) -> np.ndarray:
    """The call the reference makes: scipy KD-tree over float64 copies (Selector.py:141,243)."""
    from scipy.spatial import cKDTree

    kd = cKDTree(np.asarray(tree, dtype=np.float64))
    _, idx = kd.query(np.asarray(query, dtype=np.float64), k=k, workers=workers)
    return np.asarray(idx, dtype=np.int64).reshape(len(query), k)


def knn_graph_noself(pos: np.ndarray, k: int, use_kdtree: bool = False) -> np.ndarray:
    """torch_cluster.knn_graph(pos, k) neighbour table (GraphBuilder.py:63): k nearest OTHER points."""
    nn = (knn_kdtree if use_kdtree else knn_bruteforce)(pos, pos, k + 1)
    n = len(pos)
    out = np.empty((n, k), dtype=np.int64)
    for r in range(n):
        row = nn[r]
        keep = row[row != r]
        out[r] = keep[:k]
    return out


def sqdist_rows_fp64(tree: np.ndarray, query: np.ndarray, idx: np.ndarray) -> np.ndarray:
    t = np.asarray(tree, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    d = t[idx] - q[:, None, :]
    return (d[..., 0] ** 2 + d[..., 1] ** 2) + d[..., 2] ** 2


def tie_groups_equal(tree, query, idx_a, idx_b) -> bool:
    """True when two neighbour tables agree up to permutations inside exact-distance tie groups
    (how a SciPy result is compared: its order among exact ties is traversal-dependent)."""
    da = sqdist_rows_fp64(tree, query, idx_a)
    db = sqdist_rows_fp64(tree, query, idx_b)
    if not np.array_equal(da, db):
        return False
    diff_rows = np.nonzero((idx_a != idx_b).any(axis=1))[0]
    for r in diff_rows:
        for d in np.unique(da[r]):
            sa = np.sort(idx_a[r][da[r] == d])
            sb = np.sort(idx_b[r][db[r] == d])
            last = d == da[r].max()
            if not last and not np.array_equal(sa, sb):
                return False
    return True


# ------------------------------------------------------------------------------------------------
# small helpers reproducing torch-CPU rounding where a 0/1 decision hangs on it
# ------------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    # a*b is exact in fp64 for fp32 inputs; one rounding to fp64 then to fp32 (double rounding is
    # possible in principle, probability ~2^-29 per op)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def norm3(v: np.ndarray) -> np.ndarray:
    """torch.linalg.vector_norm over a trailing dim of 3 on CPU = sqrt(fma(z,z,fma(y,y,x*x)))."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    return np.sqrt(_fma32(z, z, _fma32(y, y, (x * x).astype(F32)))).astype(F32)


def dot3(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    p = (a * b).astype(F32)
    return ((p[..., 0] + p[..., 1]).astype(F32) + p[..., 2]).astype(F32)


def acos_threshold(rho: float) -> np.float32:
    """Largest fp32 x with torch.acos(x) > rho on this host (Decompositionor.py:290 compares acos(|x|)
    with rho); the monotone threshold form lets the kernel avoid acosf altogether."""
    import torch

    def passes(x):
        return bool((torch.tensor([x], dtype=torch.float32).acos() > rho).item())

    lo, hi = F32(0.0), F32(1.0)
    if not passes(float(lo)):
        return F32(-1.0)
    if passes(float(hi)):
        return F32(1.0)
    while np.nextafter(lo, F32(2.0)) < hi:
        mid = F32((np.float64(lo) + np.float64(hi)) / 2)
        if passes(float(mid)):
            lo = mid
        else:
            hi = mid
    return lo


def eigh3(T: np.ndarray):
    """LAPACK symmetric eigen-decomposition of a stack of 3x3 tensors, as the reference gets it."""
    import torch

    w, V = torch.linalg.eigh(torch.from_numpy(np.ascontiguousarray(T, dtype=F32)))
    return w.numpy(), V.numpy()


def _seq_sum(x: np.ndarray) -> np.ndarray:
    """fp32 sum over axis 1 in neighbour order (CPU scatter_add is sequential)."""
    acc = np.zeros((x.shape[0],) + x.shape[2:], dtype=F32)
    for a in range(x.shape[1]):
        acc = (acc + x[:, a]).astype(F32)
    return acc


# ------------------------------------------------------------------------------------------------
# PCA normals (GraphBuilder.getPVTDecompositionWithKNN, GraphBuilder.py:99-111)
# ------------------------------------------------------------------------------------------------
def pca_normals(pos: np.ndarray, nbr: np.ndarray):
    pos = np.asarray(pos, dtype=F32)
    vj = pos[nbr]                                  # [n,k,3]
    c = vj.mean(axis=1, dtype=F32)
    d = (vj - c[:, None, :]).astype(F32)
    C = (d[:, :, :, None] * d[:, :, None, :]).astype(F32).sum(axis=1, dtype=F32)
    w, V = eigh3(C)
    return V[:, :, 0].copy(), w, V


# ------------------------------------------------------------------------------------------------
# filtered normal voting tensor (Decompositionor.getBetterFilteredNVT, Decompositionor.py:278-300)
# ------------------------------------------------------------------------------------------------
def nvt_weights(pos, nrm, rows, nbr, x_thresh):
    vi = pos[rows][:, None, :]
    vj = pos[nbr]
    nj = nrm[nbr]
    dv = (vj - vi).astype(F32)
    den = np.maximum(norm3(dv), F32(1e-12))
    u = (dv / den[..., None]).astype(F32)
    x = np.abs(np.clip(dot3(u, nj), F32(-1), F32(1)))
    w = x <= x_thresh
    none = ~w.any(axis=1)
    w[none] = True                                   # :293-296
    return w


def nvt_tensor(pos, nrm, rows, nbr, x_thresh):
    pos = np.asarray(pos, dtype=F32)
    nrm = np.asarray(nrm, dtype=F32)
    w = nvt_weights(pos, nrm, rows, nbr, x_thresh)
    nj = nrm[nbr]
    outer = (nj[:, :, :, None] * nj[:, :, None, :]).astype(F32) * w[:, :, None, None].astype(F32)
    T = _seq_sum(outer) / w.sum(axis=1).astype(F32)[:, None, None]
    return T.astype(F32), w.sum(axis=1)


def nvt(pos, nrm, rows, nbr, x_thresh):
    T, sw = nvt_tensor(pos, nrm, rows, nbr, x_thresh)
    w, V = eigh3(T)
    return w, V, T, sw


# ------------------------------------------------------------------------------------------------
# eigen-space smoothing (Decomposition.getVUSmoothedNormals, Decompositionor.py:92-106)
# ------------------------------------------------------------------------------------------------
def smooth_normals(eigval, eigvec, n, tau=0.3, d=3.0):
    eigval = np.asarray(eigval, dtype=F32)
    eigvec = np.asarray(eigvec, dtype=F32)
    n = np.asarray(n, dtype=F32)
    order = np.argsort(-eigval, axis=1, kind="stable")            # descending
    lam = np.take_along_axis(eigval, order, axis=1)
    E = np.take_along_axis(eigvec, order[:, None, :], axis=2)    # E[b,c,a] = eigvec[b,c,order[a]]
    keep = (lam > F32(tau)).astype(F32)                            # indexed like the ROW of E below
    s = dot3(E, n[:, None, :])                                     # s[b,c] = E[b,c,:] . n[b]
    coef = (keep * s).astype(F32)
    t = (coef[:, :, None] * E).astype(F32)                        # [b,c,a]
    add = ((t[:, 0] + t[:, 1]).astype(F32) + t[:, 2]).astype(F32)
    m = (F32(d) * n + add).astype(F32)
    return (m / norm3(m)[:, None]).astype(F32)


# ------------------------------------------------------------------------------------------------
# labels (Decomposition.getNVTFeatures / getClasses, Decompositionor.py:57-69)
# ------------------------------------------------------------------------------------------------
def nvt_features(eigval):
    eigval = np.asarray(eigval, dtype=F32)
    l1, l2, l3 = eigval[:, 2], eigval[:, 1], eigval[:, 0]
    lin = ((l2 - l3) / l1).astype(F32)
    pla = ((l1 - l2) / l1).astype(F32)
    sph = (l3 / l1).astype(F32)
    return pla, lin, sph


def classes(eigval, scale=0.2):
    pla, lin, sph = nvt_features(eigval)
    f = np.stack([(pla * F32(scale)).astype(F32), lin, sph], axis=1)
    return np.argmax(f, axis=1).astype(np.int64)


# ------------------------------------------------------------------------------------------------
# position updates (Denoiser.py)
# ------------------------------------------------------------------------------------------------
def _apply_move(vi, x, ok, alpha, dmax, strict=True):
    t = np.where(ok[:, None], x, vi).astype(F32)
    di = ((t - vi) * F32(alpha)).astype(F32)
    ln = norm3(di)
    take = ln < F32(dmax)
    return np.where(take[:, None], (vi + di).astype(F32), vi).astype(F32)


def _solve(A, b):
    """x = inv(A) b with singular systems flagged (torch.linalg.inv_ex info != 0)."""
    A64 = A.astype(np.float64)
    ok = np.ones(len(A), dtype=bool)
    x = np.zeros_like(b, dtype=np.float64)
    det = np.linalg.det(A64)
    ok &= det != 0
    good = np.nonzero(ok)[0]
    if len(good):
        x[good] = np.linalg.solve(A64[good], b[good].astype(np.float64)[..., None])[..., 0]
    ok &= np.isfinite(x).all(axis=1)
    return x.astype(F32), ok


def flat_center_delta(pos, nbr):
    """Cloud-wide scalars of flat_step (Denoiser.py:106-107) over the multiset of gathered neighbours."""
    vj = np.asarray(pos, dtype=F32)[nbr].reshape(-1, 3)
    c = vj.mean(axis=0, dtype=np.float64).astype(F32)
    delta = norm3((vj - c).astype(F32)).max()
    return c, F32(delta)


def flat_step(pos, nrm, rows, nbr, dmax, alpha, delta=None):
    """Denoiser.flat_step, Denoiser.py:90-119.  (delta: the cloud-wide radius when `nbr` is only part of the selection.)"""
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    if delta is None:
        _, delta = flat_center_delta(pos, nbr)
    vi = pos[rows]; ni = nrm[rows]
    vj = pos[nbr]; nj = nrm[nbr]
    dist = (vj - vi[:, None]).astype(F32)
    dn = (ni[:, None] - nj).astype(F32)
    d2 = F32(delta) * F32(delta)
    sim = np.exp((F32(-16) * (dn * dn).astype(F32).sum(axis=2, dtype=F32) / d2).astype(F32))
    clo = np.exp((F32(-4) * (dist * dist).astype(F32).sum(axis=2, dtype=F32) / d2).astype(F32))
    W = (sim * clo).astype(F32)
    dt = dot3(nj, dist)
    src = ((W * dt)[:, :, None] * ni[:, None, :]).astype(F32)
    summed = _seq_sum(src)
    Ws = _seq_sum(W)
    di = (summed / Ws[:, None] * F32(alpha)).astype(F32)
    ln = norm3(di)
    di[~(ln <= F32(dmax))] = 0
    return (vi + di).astype(F32)


def feature_step(pos, nrm, rows, nbr, dmax, alpha):
    """Denoiser.feature_step, Denoiser.py:174-219."""
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    vi = pos[rows]; ni = nrm[rows]
    vj = pos[nbr]; nj = nrm[nbr]
    k = nbr.shape[1]
    nio = (ni[:, :, None] * ni[:, None, :]).astype(F32)
    njo = (nj[:, :, :, None] * nj[:, :, None, :]).astype(F32)
    A = ((np.eye(3, dtype=F32)[None] + nio) + _seq_sum(njo) + F32(k) * nio).astype(F32)
    b0 = (vi + np.einsum("nij,nj->ni", nio, vi)).astype(F32)
    b1 = np.einsum("nij,nj->ni", nio, _seq_sum(vj)).astype(F32)
    b2 = _seq_sum(np.einsum("nkij,nkj->nki", njo, vj).astype(F32))
    b = ((b0 + b1) + b2).astype(F32)
    x, ok = _solve(A, b)
    return _apply_move(vi, x, ok, alpha, dmax)


def edge_step(pos, nrm, edge_vec, rows, nbr, dmax, alpha):
    """Denoiser.edge_step, Denoiser.py:53-88."""
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    vi = pos[rows]; y = np.asarray(edge_vec, dtype=F32)[rows]
    vj = pos[nbr]; nj = nrm[nbr]
    yb = y[:, None, :]
    pv = dot3((vj - vi[:, None]).astype(F32), yb)
    pn = dot3(nj, yb)
    vp = (vj - pv[..., None] * yb).astype(F32)
    npj = (nj - pn[..., None] * yb).astype(F32)
    nn = (npj[:, :, :, None] * npj[:, :, None, :]).astype(F32)
    yy = (y[:, :, None] * y[:, None, :]).astype(F32)
    A = _seq_sum((nn + yy[:, None]).astype(F32))
    rhs = (np.einsum("nkij,nkj->nki", nn, vp) + np.einsum("nij,nj->ni", yy, vi)[:, None]).astype(F32)
    b = _seq_sum(rhs)
    x, ok = _solve(A, b)
    return _apply_move(vi, x, ok, alpha, dmax)


def corner_step(pos, nrm, rows, nbr, dmax, alpha):
    """Denoiser.corner_step, Denoiser.py:26-51."""
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    vi = pos[rows]
    vj = pos[nbr]; nj = nrm[nbr]
    njo = (nj[:, :, :, None] * nj[:, :, None, :]).astype(F32)
    A = _seq_sum(njo)
    b = _seq_sum(np.einsum("nkij,nkj->nki", njo, vj).astype(F32))
    x, ok = _solve(A, b)
    return _apply_move(vi, x, ok, alpha, dmax)


def solve_condition(A):
    return np.linalg.cond(A.astype(np.float64))


# ------------------------------------------------------------------------------------------------
# metrics (TorchUtils, Utils.py:253-304)
# ------------------------------------------------------------------------------------------------
def nn_index(tree, query):
    return knn_kdtree(np.asarray(tree, dtype=F32), np.asarray(query, dtype=F32), 1, workers=-1)[:, 0]


def chamfer_distance(pos0, pos1):
    """TorchUtils.ChamferDistance, Utils.py:253-265: un-reduced [len(pos1) + len(pos0)] squared distances."""
    pos0 = np.asarray(pos0, dtype=F32); pos1 = np.asarray(pos1, dtype=F32)
    a = (pos0[nn_index(pos0, pos1)] - pos1).astype(F32)
    b = (pos1[nn_index(pos1, pos0)] - pos0).astype(F32)
    return np.concatenate([(a * a).astype(F32).sum(axis=1, dtype=F32), (b * b).astype(F32).sum(axis=1, dtype=F32)])


def single_chamfer_distance(gt, pos):
    """sCD: not defined in the reference checkout (PostProcessing.ipynb#c9 names
    TorchUtils.SingleChamferDistance); defined here as the first half of ChamferDistance(gt, pos)."""
    gt = np.asarray(gt, dtype=F32); pos = np.asarray(pos, dtype=F32)
    a = (gt[nn_index(gt, pos)] - pos).astype(F32)
    return (a * a).astype(F32).sum(axis=1, dtype=F32)


def hausdorff_distance(pos0, pos1):
    """TorchUtils.HausdorffDistance, Utils.py:267-279."""
    pos0 = np.asarray(pos0, dtype=F32); pos1 = np.asarray(pos1, dtype=F32)
    a = norm3((pos0[nn_index(pos0, pos1)] - pos1).astype(F32))
    b = norm3((pos1[nn_index(pos1, pos0)] - pos0).astype(F32))
    return np.concatenate([a, b])


def paper_distance(gt, noisy):
    """TorchUtils.PaperDistance, Utils.py:281-295."""
    gt = np.asarray(gt, dtype=F32); noisy = np.asarray(noisy, dtype=F32)
    diag = norm3((gt.max(axis=0) - gt.min(axis=0)).astype(F32)[None])[0]
    return (norm3((gt[nn_index(gt, noisy)] - noisy).astype(F32)) / diag).astype(F32)


def average_edge_length(pos, nbr, rows=None):
    """TorchUtils.averageEdgeLength, Utils.py:298-299 over the edges (row -> nbr)."""
    pos = np.asarray(pos, dtype=F32)
    rows = np.arange(len(nbr)) if rows is None else rows
    d = (pos[nbr] - pos[rows][:, None]).astype(F32)
    return F32(norm3(d).mean(dtype=np.float64))


def pointcloud_radius(pos):
    pos = np.asarray(pos, dtype=F32)
    return norm3((pos - pos.mean(axis=0, dtype=np.float64).astype(F32)).astype(F32)).max()


# ------------------------------------------------------------------------------------------------
# normal orientation (GraphBuilder.flipNormals, GraphBuilder.py:129-209)
# ------------------------------------------------------------------------------------------------
def orient_normals(pos, nrm, nbr):
    """Edge cost 1-|ni.nj| on the kNN graph, minimum spanning tree (Kruskal, edges in ascending cost,
    stable), propagate from the top-most point flipping a child when n_parent.n_child < cos(7pi/12).
    The reference's argsort is not stable, so equal-cost edges may be taken in another order there:
    the result is compared through the propagated signs, which differ only if the trees differ."""
    pos = np.asarray(pos, dtype=F32); n = np.array(nrm, dtype=F32, copy=True)
    N, k = nbr.shape
    src = np.repeat(np.arange(N), k); dst = nbr.reshape(-1)
    cost = (F32(1) - np.abs(dot3(n[src], n[dst]))).astype(F32)
    order = np.argsort(cost, kind="stable")
    parent = np.arange(N)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    adj = [[] for _ in range(N)]
    for e in order:
        a, b = int(src[e]), int(dst[e])
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            adj[a].append(b); adj[b].append(a)
    thr = math.cos(7.0 / 12.0 * math.pi)
    root = int(np.argmax(pos[:, 2]))
    if n[root, 2] < 0:
        n[root] *= -1
    seen = np.zeros(N, dtype=bool); seen[root] = True
    stack = [root]
    while stack:
        a = stack.pop()
        for b in adj[a]:
            if not seen[b]:
                seen[b] = True
                if dot3(n[a][None], n[b][None])[0] < thr:
                    n[b] *= -1
                stack.append(b)
    return n


def orient_normals_scipy(pos, nrm, edge_index):
    """The same definition through SciPy's csgraph (another tie order among equal-cost edges than the Kruskal above):
    edge_index [2, E] as the mirror's GraphBuilder.getKNNEdgeIndex returns it.  Cross-check only."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import breadth_first_order, minimum_spanning_tree
    pos = np.asarray(pos, dtype=F32); n = np.array(nrm, dtype=F32, copy=True)
    ei = np.asarray(edge_index)
    N = len(pos)
    # csgraph treats explicit zeros as missing edges: shift the costs into (0, 2]
    cost = (F32(1) - np.abs(dot3(n[ei[0]], n[ei[1]]))).astype(np.float64) + 1e-9
    lo, hi = np.minimum(ei[0], ei[1]), np.maximum(ei[0], ei[1])
    order = np.lexsort((cost, hi, lo))
    lo, hi, cost = lo[order], hi[order], cost[order]
    first = np.ones(len(lo), dtype=bool)
    first[1:] = (lo[1:] != lo[:-1]) | (hi[1:] != hi[:-1])
    mst = minimum_spanning_tree(coo_matrix((cost[first], (lo[first], hi[first])), shape=(N, N)).tocsr())
    mst = (mst + mst.T).tocsr()
    thr = math.cos(7.0 / 12.0 * math.pi)
    root = int(np.argmax(pos[:, 2]))
    if n[root, 2] < 0:
        n[root] *= -1
    visit, parent = breadth_first_order(mst, root, directed=False, return_predecessors=True)
    for v in visit[1:]:
        p = parent[v]
        if float((n[p] * n[v]).sum(dtype=np.float32)) < thr:
            n[v] *= -1
    return n


# ------------------------------------------------------------------------------------------------
# the iterate loop (Processor.getMyFeatureDecomposition :110-117, Processor.denoise :119-139)
# ------------------------------------------------------------------------------------------------
def feature_decomposition(pos, nrm, nbr_f, x_thresh, tau=0.3, damp=3.0):
    rows = np.arange(len(pos))
    w1, V1, _, _ = nvt(pos, nrm, rows, nbr_f, x_thresh)
    f = smooth_normals(w1, V1, nrm, tau, damp)
    w2, V2, _, _ = nvt(pos, f, rows, nbr_f, x_thresh)
    return w2, V2, f, (w1, V1)


def denoise_iteration(tree, pos, nrm, k_f=16, k_u=8, x_thresh=None, alphas=(1.0, 0.2, 1.0), dmax=None,
                      strategy=("flat", "edge", "feature"), knn=knn_kdtree, scale=0.2, tau=0.3, damp=3.0,
                      snapshot=False, original=None, clamp=None):
    """One body of Processor.denoise (Processor.py:124-139) against the frozen tree `tree`.
    snapshot=True + original + clamp: the notebook's variant (PostProcessing.ipynb#c9, "Ours" / "CPSD"): every class reads the
    positions the iteration started from (results go to temp_pos), and a new position is kept only where
    |temp_pos - original| < clamp."""
    pos = np.array(pos, dtype=F32, copy=True); nrm = np.asarray(nrm, dtype=F32)
    nbr_f = knn(tree, pos, k_f)
    w2, V2, f, _ = feature_decomposition(pos, nrm, nbr_f, x_thresh, tau, damp)
    lab = classes(w2, scale)
    nbr_u = knn(tree, pos, k_u)
    edge_vec = V2[:, :, 0]
    src = pos.copy() if snapshot else pos
    for key in range(3):
        rows = np.nonzero(lab == key)[0]
        if len(rows) == 0:
            continue
        kind = strategy[key]
        sub = nbr_u[rows]
        if kind == "flat":
            new = flat_step(src, f, rows, sub, dmax, alphas[key])
        elif kind == "edge":
            new = edge_step(src, f, edge_vec, rows, sub, dmax, alphas[key])
        elif kind == "feature":
            new = feature_step(src, f, rows, sub, dmax, alphas[key])
        elif kind == "corner":
            new = corner_step(src, f, rows, sub, dmax, alphas[key])
        else:
            new = src[rows]
        pos[rows] = new
    if original is not None and clamp is not None:
        keep = norm3((pos - np.asarray(original, dtype=F32)).astype(F32)) < F32(clamp)
        pos = np.where(keep[:, None], pos, src).astype(F32)
    return pos, f, lab, (w2, V2)


def denoise(tree, pos, nrm, iterations=2, k_f=16, k_u=8, rho=5 * math.pi / 12, knn=knn_kdtree, **kw):
    """Processor.denoise, Processor.py:119-139: d = 2 * mean edge length of the 6-NN selection (self
    edge included), two iterations, alphas (1, .2, 1)."""
    x_thresh = acos_threshold(rho)
    pos = np.asarray(pos, dtype=F32)
    l = average_edge_length(pos, knn(tree, pos, 6))
    d = F32(2) * l
    labels = None
    for _ in range(iterations):
        pos, nrm, labels, _ = denoise_iteration(tree, pos, nrm, k_f, k_u, x_thresh, dmax=d, knn=knn, **kw)
    return pos, nrm, labels


# ------------------------------------------------------------------------------------------------
# Yadav-2018 baseline path ("CPSD"): radius selection, normal-filtered NVT / PVT, VU labels
# ------------------------------------------------------------------------------------------------
def ball_selection(tree: np.ndarray, query: np.ndarray, radii, chunk: int = 1024):
    """Selector.getPointsInRangeSelectionVectorized, Selector.py:214-229 = scipy KDTree.query_ball_point(query, radii):
    per query every tree point with fp64 ((dx^2+dy^2)+dz^2) <= r^2 (r upcast from fp32), ascending by index (SciPy sorts
    multi-point queries).  Returns (j int64 [total], slices int64 [m+1])."""
    t = np.asarray(tree, dtype=F32).astype(np.float64)
    q = np.asarray(query, dtype=F32).astype(np.float64)
    r = np.broadcast_to(np.asarray(radii, dtype=F32), (len(q),)).astype(np.float64)
    rows = []
    for s in range(0, len(q), chunk):
        d = q[s:s + chunk, None, :] - t[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        inside = d2 <= (r[s:s + chunk] * r[s:s + chunk])[:, None]
        rows.extend(np.nonzero(row)[0] for row in inside)
    slices = np.zeros(len(q) + 1, dtype=np.int64)
    slices[1:] = np.cumsum([len(x) for x in rows])
    j = np.concatenate(rows).astype(np.int64) if rows else np.zeros(0, np.int64)
    return j, slices


def acos_threshold_le(rho: float) -> np.float32:
    """Smallest fp32 x with torch.acos(x) <= rho: `acos(clamp(ni.nj)) <= rho` <=> `clamp(ni.nj) >= x`."""
    import torch

    def passes(x):
        return bool((torch.tensor([x], dtype=torch.float32).acos() <= rho).item())

    lo, hi = F32(-1.0), F32(1.0)
    if not passes(1.0):
        return F32(2.0)
    if passes(-1.0):
        return F32(-1.0)
    while np.nextafter(lo, F32(2.0)) < hi:
        mid = F32((np.float64(lo) + np.float64(hi)) / 2)
        if passes(float(mid)):
            hi = mid
        else:
            lo = mid
    return hi


def _rows_by_length(slices):
    """ragged CSR rows grouped by length: yields (row ids, [len(rows), L] positions into j)"""
    lens = np.diff(slices)
    for L in np.unique(lens):
        rows = np.nonzero(lens == L)[0]
        yield rows, int(L), slices[rows][:, None] + np.arange(int(L))[None, :]


def nvt_normal_filtered(nrm, centres, j, slices, x_le):
    """Decompositionor.getNormalFilteredNVT, Decompositionor.py:260-276.  Returns (eigval, eigvec, T, sum w)."""
    nrm = np.asarray(nrm, dtype=F32)
    m = len(slices) - 1
    T = np.zeros((m, 3, 3), dtype=F32)
    sw = np.zeros(m, dtype=np.int64)
    for rows, L, at in _rows_by_length(slices):
        ni = nrm[centres[rows]]
        if L == 0:
            T[rows] = (ni[:, :, None] * ni[:, None, :]).astype(F32)
            continue
        nj = nrm[j[at]]
        w = np.clip(dot3(ni[:, None, :], nj), F32(-1), F32(1)) >= x_le
        outer = (nj[:, :, :, None] * nj[:, :, None, :]).astype(F32) * w[:, :, None, None].astype(F32)
        cnt = w.sum(axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            Tr = (_seq_sum(outer) / cnt.astype(F32)[:, None, None]).astype(F32)
        none = cnt == 0
        Tr[none] = (ni[none][:, :, None] * ni[none][:, None, :]).astype(F32)
        T[rows] = Tr
        sw[rows] = cnt
    w_, V_ = eigh3(T)
    return w_, V_, T, sw


def pvt_normal_filtered(pos, nrm, centres, j, slices, x_le):
    """Decompositionor.getNormalFilteredPVT, Decompositionor.py:172-211.  Returns (eigval, eigvec, C, sum w)."""
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    m = len(slices) - 1
    C = np.zeros((m, 3, 3), dtype=F32)
    sw = np.zeros(m, dtype=np.int64)
    for rows, L, at in _rows_by_length(slices):
        ni = nrm[centres[rows]]
        if L == 0:
            v = pos[centres[rows]]
            s1 = np.cross(ni, v).astype(F32)
            s2 = np.cross(ni, s1).astype(F32)
            acc = np.zeros((len(rows), 3, 3), dtype=F32)
            for smp in (s1, -s1, s2, -s2):
                acc = (acc + (smp[:, :, None] * smp[:, None, :]).astype(F32)).astype(F32)
            C[rows] = acc
            continue
        nj = nrm[j[at]]; vj = pos[j[at]]
        w = np.clip(dot3(ni[:, None, :], nj), F32(-1), F32(1)) >= x_le
        w[~w.any(axis=1)] = True                                      # :186-189
        cnt = w.sum(axis=1)
        wf = w[:, :, None].astype(F32)
        c = (_seq_sum((vj * wf).astype(F32)) / cnt.astype(F32)[:, None]).astype(F32)
        dv = (vj - c[:, None, :]).astype(F32)
        outer = (dv[:, :, :, None] * dv[:, :, None, :]).astype(F32) * w[:, :, None, None].astype(F32)
        C[rows] = (_seq_sum(outer) / cnt.astype(F32)[:, None, None]).astype(F32)
        sw[rows] = cnt
    w_, V_ = eigh3(C)
    return w_, V_, C, sw


def vu_features(eigval, tau):
    """Decomposition.getVUFeatures, Decompositionor.py:84-85."""
    return ((np.asarray(eigval, dtype=F32) < F32(tau)).sum(axis=1) % 3).astype(np.int64)


def csr_step(kind, pos, nrm, centres, j, slices, dmax, alpha, edge_vec=None):
    """One Denoiser step on ragged rows (Selection.filter of a radius selection): dense step per row length; flat_step's
    cloud-wide centre / delta (Denoiser.py:106-107) is taken over ALL rows first."""
    pos = np.asarray(pos, dtype=F32)
    out = np.zeros((len(slices) - 1, 3), dtype=F32)
    delta = None
    if kind == "flat":
        _, delta = flat_center_delta(pos, j)
    for rows, L, at in _rows_by_length(slices):
        c, nb = centres[rows], j[at]
        if L == 0:
            out[rows] = pos[c]
            continue
        if kind == "flat":
            out[rows] = flat_step(pos, nrm, c, nb, dmax, alpha, delta=delta)
        elif kind == "edge":
            out[rows] = edge_step(pos, nrm, edge_vec, c, nb, dmax, alpha)
        elif kind == "corner":
            out[rows] = corner_step(pos, nrm, c, nb, dmax, alpha)
        else:
            out[rows] = feature_step(pos, nrm, c, nb, dmax, alpha)
    return out


def cpsd_iteration(tree, pos, nrm, original_pos, d, rho=0.9, tau=0.3, alphas=(0.1, 1.0, 1.0), k_u=8, knn=None):
    """One iteration of the notebook's "CPSD" loop (PostProcessing.ipynb#c9, j == 1): Martin feature decomposition at radius d,
    VU labels, the three class steps computed from the SAME snapshot (unlike Processor.denoise), accepted where the point
    stays within d of its original position.  Returns (new positions, smoothed normals, labels, temp positions)."""
    knn = knn or knn_kdtree
    pos = np.asarray(pos, dtype=F32); nrm = np.asarray(nrm, dtype=F32)
    n = len(pos)
    centres = np.arange(n)
    j, slices = ball_selection(tree, pos, np.full(n, F32(d), dtype=F32))
    x_le = acos_threshold_le(rho)
    w1, V1, _, _ = nvt_normal_filtered(nrm, centres, j, slices, x_le)
    f_n = smooth_normals(w1, V1, nrm)
    w2, V2, _, _ = pvt_normal_filtered(pos, f_n, centres, j, slices, x_le)
    lab = vu_features(w2, tau)
    nb8 = knn(tree, pos, k_u)
    temp = pos.copy()
    big = F32(d) * F32(20000)
    for key in range(3):
        rows = np.nonzero(lab == key)[0]
        if len(rows) == 0:
            continue
        if key == 0:
            temp[rows] = flat_step(pos, f_n, rows, nb8[rows], big, alphas[0])
        elif key == 1:
            temp[rows] = edge_step(pos, f_n, V2[:, :, 0], rows, nb8[rows], big, alphas[1])
        else:
            temp[rows] = corner_step(pos, f_n, rows, nb8[rows], big, alphas[2])
    mask = norm3((temp - np.asarray(original_pos, dtype=F32)).astype(F32)) < F32(d)
    new = pos.copy()
    new[mask] = temp[mask]
    return new, f_n, lab, temp


# ------------------------------------------------------------------------------------------------
# mesh vertex update (PatchGeneration.Modules.Mesh.updateVertices, Mesh.py:377-418)
# ------------------------------------------------------------------------------------------------
def mesh_vertex_update(v, f, face_normals, vta_faces, vta_offsets, k=15):
    """k Jacobi sweeps of  v_i += 1/(3 deg_i) sum_{f at i} sum_{c in f} n_f (n_f . (v_c - v_i)), fp64, every vertex from
    the same snapshot.  Sums run over the incident faces first (per corner), then over the three corners, as the
    reference's two np.sum calls do."""
    v = np.array(v, dtype=np.float64)
    f = np.asarray(f, dtype=np.int64)
    n = np.asarray(face_normals, dtype=np.float64)
    vf = np.asarray(vta_faces, dtype=np.int64)
    ni = np.asarray(vta_offsets, dtype=np.int64)
    deg = np.diff(ni)
    owner = np.repeat(np.arange(len(v)), deg)
    for _ in range(k):
        nj = n[vf]                                        # [E,3]
        dvs = v[f[vf]] - v[owner][:, None, :]             # [E,3 corners,3]
        dot = ((nj[:, None, 0] * dvs[:, :, 0] + nj[:, None, 1] * dvs[:, :, 1]) + nj[:, None, 2] * dvs[:, :, 2])
        el = dot[:, :, None] * nj[:, None, :]             # [E,corner,axis]
        S = np.zeros((len(v), 3, 3))
        np.add.at(S, owner, el)                           # sequential over the incident faces, in adjacency order
        S = (S[:, 0] + S[:, 1]) + S[:, 2]
        with np.errstate(invalid="ignore", divide="ignore"):
            v = v + S / (3 * deg)[:, None]
    return v
