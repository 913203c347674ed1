"""ctypes binding of libngpd.so (include/ngpd.h).  Thin: pointer extraction, argument checks, error
propagation.  There is no CPU implementation behind any of these calls: tensors must live on a CUDA device
and the shared library must have been built (python __graft_entry__.py / build.py)."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGPD_LIBRARY") or os.path.join(_HERE, "libngpd.so")   # NGPD_LIBRARY: another build of the same library (A/B timing)

c_i64, c_i32, c_f32, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p

KNN_SKIP_SELF, KNN_QUERY_IS_TREE, KNN_COHERENT, KNN_EXACT_ONLY = 1, 2, 4, 8
STEP_SNAPSHOT_CLASSES = 1
STEP_FLAT, STEP_EDGE, STEP_FEATURE, STEP_CORNER, STEP_NONE = 0, 1, 2, 3, -1


class GridInfo(ctypes.Structure):
    _fields_ = [("n", c_i64), ("cell_size", ctypes.c_double), ("dims", c_i32 * 3), ("bricks", c_i32),
                ("occupied_cells", c_i64), ("bytes", c_i64), ("bbox", c_f32 * 6), ("rebuilds", c_i32)]


class StepParams(ctypes.Structure):
    _fields_ = [("k_feature", c_i32), ("k_update", c_i32), ("x_thresh", c_f32), ("tau", c_f32), ("damp", c_f32),
                ("scale", c_f32), ("strategy", c_i32 * 3), ("alpha", c_f32 * 3), ("dmax", c_f32), ("flags", c_i32),
                ("clamp_radius", c_f32)]


class SlabWiring(ctypes.Structure):
    _fields_ = [("world", c_i32), ("rank", c_i32), ("n_send", c_i64), ("n_recv", c_i64), ("cap", c_i64), ("send_rows", c_vp),
                ("recv_rows", c_vp), ("send_seg_host", ctypes.POINTER(c_i64)), ("first_row_host", ctypes.POINTER(c_i64)),
                ("symm_base_host", ctypes.POINTER(ctypes.c_uint64))]


# every symbol declared in include/ngpd.h, with its ctypes signature
SIGNATURES = {
    "ngpd_last_error": (ctypes.c_char_p, []),
    "ngpd_version": (ctypes.c_int, []),
    "ngpd_grid_create": (ctypes.c_int, [c_vp, c_i64, c_f32, ctypes.c_int, c_vp, ctypes.POINTER(c_vp)]),
    "ngpd_grid_destroy": (ctypes.c_int, [c_vp]),
    "ngpd_grid_info": (ctypes.c_int, [c_vp, ctypes.POINTER(GridInfo)]),
    "ngpd_grid_order": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "ngpd_knn": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp]),
    "ngpd_nn_sqdist": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_vp, c_vp]),
    "ngpd_nn_sqdist_reduce": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_ball_query": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_nvt_normal": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_pvt_normal": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_mesh_vertex_update": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp]),
    "ngpd_orient_normals": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_f32, ctypes.POINTER(c_i32), c_vp]),
    "ngpd_pca_normals": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_nvt": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_eigh3": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "ngpd_smooth_normals": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "ngpd_classify": (ctypes.c_int, [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "ngpd_center_delta": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "ngpd_update": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_f32, c_f32, c_vp, c_vp, c_vp]),
    "ngpd_edge_length_sum": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_vp]),
    "ngpd_session_create": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, c_vp, ctypes.POINTER(c_vp)]),
    "ngpd_session_destroy": (ctypes.c_int, [c_vp]),
    "ngpd_session_reserve": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "ngpd_session_set_state": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ngpd_session_get_state": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_session_set_owned": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "ngpd_session_step": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), c_vp]),
    "ngpd_session_phase_features": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), ctypes.c_int, c_vp]),
    "ngpd_session_phase_flat_scalars": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), ctypes.c_int, ctypes.c_int, c_vp]),
    "ngpd_session_phase_update": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), ctypes.c_int, c_vp]),
    "ngpd_session_phase_commit_normals": (ctypes.c_int, [c_vp]),
    "ngpd_session_mean_edge_length": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), c_vp]),
    "ngpd_session_launch_count": (ctypes.c_int, [c_vp]),
    "ngpd_session_set_knn_mode": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "ngpd_session_last_fixups": (ctypes.c_int, [c_vp, c_vp]),
    "ngpd_session_knn_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i32), c_vp]),
    "ngpd_session_set_profiling": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "ngpd_session_get_profile": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i32)]),
    "ngpd_session_order": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "ngpd_session_buffer": (c_vp, [c_vp, ctypes.c_int]),
    "ngpd_session_export_rows": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64, c_vp, c_vp]),
    "ngpd_session_export_rows_peers": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64, c_vp, c_vp, ctypes.c_int, c_vp]),
    "ngpd_session_import_rows": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64, c_vp, c_vp]),
    "ngpd_slab_symm_bytes": (c_i64, [ctypes.c_int, c_i64]),
    "ngpd_session_set_slab": (ctypes.c_int, [c_vp, ctypes.POINTER(SlabWiring), c_vp]),
    "ngpd_session_slab_refresh": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "ngpd_session_slab_allreduce": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "ngpd_session_step_slab": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), c_vp]),
    "ngpd_session_halo_need": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_f32), c_vp]),
    "ngpd_session_checksum": (ctypes.c_int, [c_vp, c_vp, ctypes.POINTER(ctypes.c_uint64), c_vp]),
    "ngpd_session_set_original": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "ngpd_session_run_host": (ctypes.c_int, [c_vp, ctypes.POINTER(StepParams), ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ngpd_denoise_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(StepParams), ctypes.c_int, c_vp, c_vp, c_vp]),
}

_lib = None


def load():
    """Load libngpd.so once.  Raises if it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                              "(nvcc, sm_100a). This package has no CPU path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class NgpdError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ngpd_last_error()
        raise NgpdError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("normal-guided-pointcloud-denoiser_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback.")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dev(t: torch.Tensor, dtype=None, name="tensor") -> torch.Tensor:
    """A contiguous CUDA tensor of the given dtype (copying only when needed)."""
    if not torch.is_tensor(t):
        raise ValueError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} lives on {t.device}; this package only computes on CUDA tensors")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


class Grid:
    """Frozen spatial index over a copy of `pos` (ngpd_grid_t)."""

    def __init__(self, pos: torch.Tensor, k_hint: int = 16, cell_size: float = 0.0):
        require_cuda()
        pos = dev(pos, torch.float32, "pos")
        assert pos.dim() == 2 and pos.size(1) == 3
        self.n = pos.size(0)
        self.device = pos.device
        h = c_vp()
        with torch.cuda.device(pos.device):
            check(load().ngpd_grid_create(ptr(pos), self.n, float(cell_size), int(k_hint), stream(), ctypes.byref(h)),
                  "ngpd_grid_create")
        self._h = h

    def info(self) -> GridInfo:
        gi = GridInfo()
        check(load().ngpd_grid_info(self._h, ctypes.byref(gi)), "ngpd_grid_info")
        return gi

    def order(self) -> torch.Tensor:
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        check(load().ngpd_grid_order(self._h, ptr(out), stream()), "ngpd_grid_order")
        return out

    def knn(self, query: torch.Tensor, k: int, flags: int = 0, with_d2: bool = False):
        query = dev(query, torch.float32, "query")
        m = query.size(0)
        idx = torch.empty((m, k), dtype=torch.int32, device=query.device)
        d2 = torch.empty((m, k), dtype=torch.float32, device=query.device) if with_d2 else None
        with torch.cuda.device(query.device):
            check(load().ngpd_knn(self._h, ptr(query), m, k, flags, ptr(idx), ptr(d2), stream()), "ngpd_knn")
        return (idx, d2) if with_d2 else idx

    def ball(self, query: torch.Tensor, radii: torch.Tensor, flags: int = 0):
        """(idx int32 [total], offsets int32 [m+1]): per query every tree point within its radius, ascending by index"""
        q = dev(query, torch.float32, "query")
        r = dev(radii, torch.float32, "radii")
        m = q.size(0)
        assert r.dim() == 1 and r.size(0) == m
        with torch.cuda.device(q.device):
            counts = torch.empty(m, dtype=torch.int32, device=q.device)
            check(load().ngpd_ball_query(self._h, ptr(q), m, ptr(r), flags, ptr(counts), None, None, stream()), "ngpd_ball_query")
            total = int(counts.sum(dtype=torch.int64).item())
            if total >= 2 ** 31:
                raise NgpdError(f"ngpd_ball_query: {total} neighbours do not fit the int32 offsets of the neighbourhood kernels")
            offsets = torch.zeros(m + 1, dtype=torch.int32, device=q.device)
            offsets[1:] = torch.cumsum(counts, 0, dtype=torch.int64).to(torch.int32)
            idx = torch.empty(max(total, 1), dtype=torch.int32, device=q.device)
            check(load().ngpd_ball_query(self._h, ptr(q), m, ptr(r), flags, None, ptr(offsets), ptr(idx), stream()), "ngpd_ball_query")
        return idx[:total], offsets

    def nn_sqdist(self, query: torch.Tensor, want_idx: bool = False, flags: int = 0):
        query = dev(query, torch.float32, "query")
        m = query.size(0)
        d2 = torch.empty(m, dtype=torch.float32, device=query.device)
        idx = torch.empty(m, dtype=torch.int32, device=query.device) if want_idx else None
        with torch.cuda.device(query.device):
            check(load().ngpd_nn_sqdist(self._h, ptr(query), m, flags, ptr(d2), ptr(idx), stream()), "ngpd_nn_sqdist")
        return (d2, idx) if want_idx else d2

    def nn_reduce(self, query: torch.Tensor, flags: int = 0, want_d2: bool = False):
        """{sum d2, sum d, max d2, rows} (fp64 device tensor [4]) of the nearest-neighbour pass, optionally with the per-query
        squared distances: the fused form of `metric(...).mean()` / `.max()`"""
        query = dev(query, torch.float32, "query")
        m = query.size(0)
        acc = torch.empty(4, dtype=torch.float64, device=query.device)
        d2 = torch.empty(m, dtype=torch.float32, device=query.device) if want_d2 else None
        with torch.cuda.device(query.device):
            check(load().ngpd_nn_sqdist_reduce(self._h, ptr(query), m, flags, ptr(d2), None, ptr(acc), stream()), "ngpd_nn_sqdist_reduce")
        return (acc, d2) if want_d2 else acc

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ngpd_grid_destroy(h)


_THRESH_CACHE: dict = {}


def acos_threshold(rho: float) -> float:
    """Largest fp32 x with torch-CPU acos(x) > rho: the form in which the kernels evaluate the reference's
    `acos(|u.n|) > rho` neighbour filter (Decompositionor.py:290) without calling acosf."""
    rho = float(rho)
    if rho not in _THRESH_CACHE:
        import numpy as np

        def passes(x):
            return bool((torch.tensor([x], dtype=torch.float32).acos() > rho).item())

        lo, hi = np.float32(0.0), np.float32(1.0)
        if not passes(0.0):
            val = -1.0
        elif passes(1.0):
            val = 1.0
        else:
            while np.nextafter(lo, np.float32(2.0)) < hi:
                mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
                if passes(float(mid)):
                    lo = mid
                else:
                    hi = mid
            val = float(lo)
        _THRESH_CACHE[rho] = val
    return _THRESH_CACHE[rho]


def acos_threshold_le(rho: float) -> float:
    """Smallest fp32 x in [-1, 1] with torch-CPU acos(x) <= rho: `acos(clamp(ni.nj)) <= rho` (Decompositionor.py:185,
    271) becomes `clamp(ni.nj) >= x`.  2.0 when nothing passes (rho < 0)."""
    key = ("le", float(rho))
    if key not in _THRESH_CACHE:
        import numpy as np

        def passes(x):
            return bool((torch.tensor([x], dtype=torch.float32).acos() <= float(rho)).item())

        lo, hi = np.float32(-1.0), np.float32(1.0)       # passes(hi) is acos(1) = 0 <= rho
        if not passes(1.0):
            val = 2.0
        elif passes(-1.0):
            val = -1.0
        else:
            while np.nextafter(lo, np.float32(2.0)) < hi:
                mid = np.float32((np.float64(lo) + np.float64(hi)) / 2)
                if passes(float(mid)):
                    hi = mid
                else:
                    lo = mid
            val = float(hi)
        _THRESH_CACHE[key] = val
    return _THRESH_CACHE[key]


class Session:
    """Fused tree-order denoising state (ngpd_session_t)."""

    def __init__(self, tree_pos: torch.Tensor, k_hint: int = 16):
        require_cuda()
        tree_pos = dev(tree_pos, torch.float32, "tree_pos")
        self.n = tree_pos.size(0)
        self.device = tree_pos.device
        h = c_vp()
        with torch.cuda.device(self.device):
            check(load().ngpd_session_create(ptr(tree_pos), self.n, int(k_hint), stream(), ctypes.byref(h)),
                  "ngpd_session_create")
        self._h = h

    def reserve(self, k_feature: int):
        """allocate the step's working buffers now instead of inside the first iteration"""
        with torch.cuda.device(self.device):
            check(load().ngpd_session_reserve(self._h, int(k_feature), stream()), "ngpd_session_reserve")

    def set_state(self, pos, nrm):
        pos = dev(pos, torch.float32, "pos") if pos is not None else None
        nrm = dev(nrm, torch.float32, "nrm") if nrm is not None else None
        with torch.cuda.device(self.device):
            check(load().ngpd_session_set_state(self._h, ptr(pos), ptr(nrm), stream()), "ngpd_session_set_state")

    def get_state(self, want_labels=True):
        pos = torch.empty((self.n, 3), dtype=torch.float32, device=self.device)
        nrm = torch.empty((self.n, 3), dtype=torch.float32, device=self.device)
        lab = torch.empty(self.n, dtype=torch.uint8, device=self.device) if want_labels else None
        with torch.cuda.device(self.device):
            check(load().ngpd_session_get_state(self._h, ptr(pos), ptr(nrm), ptr(lab), stream()), "ngpd_session_get_state")
        return pos, nrm, lab

    def step(self, params: StepParams):
        with torch.cuda.device(self.device):
            check(load().ngpd_session_step(self._h, ctypes.byref(params), stream()), "ngpd_session_step")

    def mean_edge_length_parts(self, k: int):
        out = (ctypes.c_double * 2)()
        with torch.cuda.device(self.device):
            check(load().ngpd_session_mean_edge_length(self._h, k, out, stream()), "ngpd_session_mean_edge_length")
        return out[0], out[1]

    PROFILE_NAMES = ("knn", "nvt_smooth", "nvt_classify", "flat_scalars", "update", "halo")

    def set_profiling(self, on: bool):
        check(load().ngpd_session_set_profiling(self._h, int(on)), "ngpd_session_set_profiling")

    def get_profile(self) -> dict:
        ms = (ctypes.c_double * len(self.PROFILE_NAMES))()
        cnt = (c_i32 * len(self.PROFILE_NAMES))()
        check(load().ngpd_session_get_profile(self._h, ms, cnt), "ngpd_session_get_profile")
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROFILE_NAMES)}

    CHECKSUM_FIELDS = ("pos_hash", "nrm_hash", "sum_x_q24", "sum_y_q24", "sum_z_q24", "sum_norm_q24", "label0", "label1", "label2",
                       "label_other", "rows")

    def checksum(self, global_ids: torch.Tensor | None = None) -> list[int]:
        """Order- and partition-independent digest of the owned rows (integers; ngpd_session_checksum).  Digests of
        disjoint slabs add up (mod 2^64) to the digest of the whole cloud."""
        out = (ctypes.c_uint64 * len(self.CHECKSUM_FIELDS))()
        if global_ids is not None:
            global_ids = dev(global_ids, torch.int64, "global_ids")
            assert global_ids.numel() == self.n
        with torch.cuda.device(self.device):
            check(load().ngpd_session_checksum(self._h, ptr(global_ids), out, stream()), "ngpd_session_checksum")
        return [int(v) for v in out]

    def halo_need(self, reset: bool = False) -> float:
        out = c_f32()
        with torch.cuda.device(self.device):
            check(load().ngpd_session_halo_need(self._h, int(reset), ctypes.byref(out), stream()), "ngpd_session_halo_need")
        return float(out.value)

    def set_original(self, pos: torch.Tensor | None):
        """positions the displacement clamp (StepParams.clamp_radius) is measured from, original point order"""
        pos = dev(pos, torch.float32, "pos") if pos is not None else None
        with torch.cuda.device(self.device):
            check(load().ngpd_session_set_original(self._h, ptr(pos), stream()), "ngpd_session_set_original")

    def set_knn_mode(self, mode):
        """0 / False: all tiers (default); 1 / True: exact shell search only; 2: streaming tiers without re-ranking"""
        check(load().ngpd_session_set_knn_mode(self._h, int(mode)), "ngpd_session_set_knn_mode")

    def last_fixups(self) -> int:
        return load().ngpd_session_last_fixups(self._h, stream())

    def knn_stats(self):
        """(rows tier 0 handed to the search, rows the 3x3x3 tier handed on, rows the 5x5x5 tier handed to the exact
        search) of the last kNN pass"""
        out = (c_i32 * 3)()
        check(load().ngpd_session_knn_stats(self._h, out, stream()), "ngpd_session_knn_stats")
        return int(out[0]), int(out[1]), int(out[2])

    def launch_count(self) -> int:
        return load().ngpd_session_launch_count(self._h)

    def order(self) -> torch.Tensor:
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        check(load().ngpd_session_order(self._h, ptr(out), stream()), "ngpd_session_order")
        return out

    def run_host(self, params: StepParams, iterations: int, pos_host, nrm_host, pos_out, nrm_out, labels_out):
        for t in (pos_host, nrm_host, pos_out, nrm_out, labels_out):
            assert t is None or (not t.is_cuda and t.is_contiguous())
        with torch.cuda.device(self.device):
            check(load().ngpd_session_run_host(self._h, ctypes.byref(params), iterations, ptr(pos_host), ptr(nrm_host),
                                               ptr(pos_out), ptr(nrm_out), ptr(labels_out), stream()), "ngpd_session_run_host")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ngpd_session_destroy(h)


def make_params(k_feature=16, k_update=8, rho=None, tau=0.3, damp=3.0, scale=0.2,
                strategy=(STEP_FLAT, STEP_EDGE, STEP_FEATURE), alpha=(1.0, 0.2, 1.0), dmax=0.0, flags=0,
                clamp_radius=0.0) -> StepParams:
    import math

    p = StepParams()
    p.k_feature, p.k_update = int(k_feature), int(k_update)
    p.x_thresh = acos_threshold(math.pi * 5 / 12 if rho is None else rho)
    p.tau, p.damp, p.scale = float(tau), float(damp), float(scale)
    for i in range(3):
        p.strategy[i] = int(strategy[i])
        p.alpha[i] = float(alpha[i])
    p.dmax = float(dmax)
    p.flags, p.clamp_radius = int(flags), float(clamp_radius)
    return p
