"""Multi-GPU partitioning of one cloud: Morton slabs + halo exchange (new work, the reference is
single-process).

The index is frozen on the construction-time positions, so ownership and halo MEMBERSHIP are fixed for a whole run
and only halo VALUES travel:
  * plan   - sort the construction-time positions by a 63-bit Morton key, cut the order into `world` contiguous
             slabs of equal size; a rank's halo is every foreign point inside the 27-neighbourhood of a coarse cell
             (edge = halo width) that holds one of its own points, i.e. every foreign point within the halo width
             of the slab.  Each rank computes its own slab and halo from the replicated input: no communication.
  * wiring - every rank tells each owner which of its points it needs (one id list per peer, exchanged once).
  * step   - per dependency phase: smoothed normals after the first tensor pass, positions after every class update
             (class-sequential semantics, Processor.py:127-138), plus two scalar all-reduces for flat_step's
             cloud-wide centre / radius (Denoiser.py:106-107).  Payload: float4 rows packed by a gather kernel,
             point-to-point over NCCL (NVLink 5 / NVSwitch); gloo on CPU in the tests.
`SlabPlan` and `HaloExchanger` are pure torch + torch.distributed (device-agnostic, covered by world_size-2 gloo tests);
`SlabSession` binds them to the CUDA session."""
from __future__ import annotations

import contextlib
import math
import os

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------
# planning (pure torch)
# ------------------------------------------------------------------------------------------------------
def _spread3(v: torch.Tensor) -> torch.Tensor:
    v = v & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton_keys(pos: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, bits: int = 21) -> torch.Tensor:
    """63-bit Morton code of the positions quantised to `bits` bits per axis over the box [lo, hi]."""
    ext = (hi - lo).max().clamp_min(1e-30).double()
    q = ((pos.double() - lo.double()) / ext * (2 ** bits - 1)).floor_().clamp_(0, 2 ** bits - 1).long()
    return _spread3(q[:, 0]) | (_spread3(q[:, 1]) << 1) | (_spread3(q[:, 2]) << 2)


class SlabPlan:
    """Ownership and halo membership of one rank.  All index tensors hold ORIGINAL point ids unless named local."""

    def __init__(self, tree_pos: torch.Tensor, rank: int, world: int, halo_width: float):
        n = tree_pos.size(0)
        dev = tree_pos.device
        self.rank, self.world, self.n, self.halo_width = rank, world, n, float(halo_width)
        lo, hi = tree_pos.min(dim=0).values, tree_pos.max(dim=0).values
        keys = morton_keys(tree_pos, lo, hi)
        order = torch.argsort(keys, stable=True)
        del keys
        bounds = [(r * n) // world for r in range(world + 1)]
        self.bounds = bounds
        slab_of = torch.empty(n, dtype=torch.int16, device=dev)
        slab_of[order] = torch.bucketize(torch.arange(n, device=dev), torch.tensor(bounds[1:-1], device=dev), right=True).to(torch.int16)
        self.slab_of = slab_of
        self.owned = order[bounds[rank]:bounds[rank + 1]].clone()
        del order
        # halo: foreign points in the 27-neighbourhood (coarse cells of edge halo_width) of an owned point's cell
        w = self.halo_width
        cell = ((tree_pos.double() - lo.double()) / w).floor_().long() + 1          # +1: room for the -1 offsets
        dims = cell.max(dim=0).values + 2
        lin = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
        mine = torch.unique(lin[self.owned])
        offs = torch.tensor([(dz * int(dims[1]) + dy) * int(dims[0]) + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)],
                            device=dev)
        near = torch.unique((mine[:, None] + offs[None, :]).reshape(-1))
        if world > 1:
            cand = torch.isin(lin, near) & (slab_of != rank)
            halo = cand.nonzero().flatten()
            owner = slab_of[halo].long()
            by_owner = torch.argsort(owner * n + halo)                            # grouped by owner, ascending id inside
            self.halo = halo[by_owner]
            self.halo_owner = owner[by_owner]
        else:
            self.halo = torch.empty(0, dtype=torch.long, device=dev)
            self.halo_owner = torch.empty(0, dtype=torch.long, device=dev)
        self.local_ids = torch.cat([self.owned, self.halo])                        # local index -> original id
        self.n_owned, self.n_halo = self.owned.numel(), self.halo.numel()
        # what I ask of each peer (ids, in the order their values will arrive)
        self.request = [self.halo[self.halo_owner == q] for q in range(world)]

    def local_index_of(self, ids: torch.Tensor) -> torch.Tensor:
        """local index (position in local_ids) of original ids that are owned by this rank"""
        look = torch.full((self.n,), -1, dtype=torch.long, device=ids.device)
        look[self.owned] = torch.arange(self.n_owned, device=ids.device)
        out = look[ids]
        assert bool((out >= 0).all()), "a peer asked for a point this rank does not own"
        return out


class HaloExchanger:
    """Fixed send/recv wiring between slabs.  `send_local[q]` = local indices (of owned points) whose values go to
    peer q; `recv_local[q]` = local indices (of halo points) that peer q fills, in matching order."""

    def __init__(self, plan: SlabPlan, group=None):
        self.plan, self.group = plan, group
        world, rank, dev = plan.world, plan.rank, plan.local_ids.device
        counts = torch.tensor([r.numel() for r in plan.request], dtype=torch.long, device=dev)
        all_counts = [torch.empty_like(counts) for _ in range(world)]
        if world > 1:
            dist.all_gather(all_counts, counts, group=group)
        else:
            all_counts = [counts]
        self.recv_counts = [int(c) for c in counts.tolist()]
        self.send_counts = [int(all_counts[q][rank]) for q in range(world)]
        wanted = [torch.empty(self.send_counts[q], dtype=torch.long, device=dev) for q in range(world)]
        self._p2p([(plan.request[q], q) for q in range(world) if q != rank and self.recv_counts[q] > 0],
                  [(wanted[q], q) for q in range(world) if q != rank and self.send_counts[q] > 0])
        self.send_local = [plan.local_index_of(wanted[q]) if self.send_counts[q] else wanted[q] for q in range(world)]
        base = plan.n_owned
        self.recv_local, off = [], 0
        for q in range(world):
            self.recv_local.append(torch.arange(base + off, base + off + self.recv_counts[q], device=dev))
            off += self.recv_counts[q]

    def _p2p(self, sends, recvs):
        ops = [dist.P2POp(dist.isend, t.contiguous(), q, group=self.group) for t, q in sends]
        ops += [dist.P2POp(dist.irecv, t, q, group=self.group) for t, q in recvs]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange(self, send_buf: torch.Tensor, recv_buf: torch.Tensor):
        """send_buf rows are ordered by peer as in send_local (concatenated), recv_buf likewise for recv_local.
        One variable-split all-to-all per exchange: a single collective call (one NCCL kernel) instead of up to
        2*(world-1) point-to-point operations -- at 8 GPUs the host-side cost of building and launching those was larger
        than the iteration's kernels (measured: 3.7 ms of a 7.7 ms step)."""
        if self.plan.world == 1:
            return
        dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts,
                               group=self.group)

    # -- halo push over peer memory (opt-in: SlabSession(transport="peer")) ---------------------------------------------
    def enable_peer_push(self, device: torch.device):
        """Symmetric receive buffers (two, used alternately) that the peers' export kernels write into directly
        (`ngpd_session_export_rows_peers`): one kernel + one cross-rank barrier per exchange instead of gather kernel ->
        NCCL all-to-all -> scatter.  Double buffering: a peer may already be pushing exchange t+1 while this rank still
        unpacks exchange t (the barrier only orders a rank's unpacking after everybody's pushing)."""
        import torch.distributed._symmetric_memory as symm
        world, rank = self.plan.world, self.plan.rank
        group = self.group if self.group is not None else dist.group.WORLD
        mine = torch.tensor(self.recv_counts, dtype=torch.long, device=device)
        table = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(table, mine, group=group)
        recv = [[int(v) for v in t.tolist()] for t in table]              # recv[q][s] = rows rank q receives from rank s
        cap, first_row, seg = peer_push_layout(recv, rank)
        assert seg[1:] == [sum(self.send_counts[:q + 1]) for q in range(world)]
        self._sym = symm.empty((2, cap, 4), dtype=torch.float32, device=device)
        self._hdl = symm.rendezvous(self._sym, group)
        ptrs = [int(a) for a in self._hdl.buffer_ptrs]
        self._peer_base = [torch.tensor([ptrs[q] + 16 * (b * cap + first_row[q]) for q in range(world)],
                                        dtype=torch.int64, device=device) for b in range(2)]
        self._seg = torch.tensor(seg, dtype=torch.long, device=device)
        self._parity = 0
        self.peer = True

    def bytes_per_exchange(self, row_bytes: int = 16):
        return sum(self.send_counts) * row_bytes, sum(self.recv_counts) * row_bytes


def peer_push_layout(recv: list[list[int]], rank: int):
    """Addressing of the peer-memory halo push.  recv[q][s] = rows rank q receives from rank s (its receive buffer holds
    them grouped by source, ascending).  Returns (rows per receive buffer -- the same on every rank, as symmetric memory
    wants --, first_row[q] = row inside rank q's buffer where THIS rank's block starts, seg = prefix sums of this rank's
    send counts: rows [seg[q], seg[q+1]) of its packed send list belong to rank q)."""
    world = len(recv)
    cap = max(1, max(sum(r) for r in recv))
    first_row = [sum(recv[q][:rank]) for q in range(world)]
    seg = [0]
    for q in range(world):
        seg.append(seg[-1] + recv[q][rank])
    return cap, first_row, seg


def estimate_halo_width(tree_pos: torch.Tensor, k: int, factor: float = 6.0) -> float:
    """A safe halo width: `factor` times the typical k-NN radius, estimated from the bounding-box surface density
    (the same first guess the grid builder uses).  The k-NN radius of a moved query must stay below
    halo_width - displacement for the slab result to equal the single-GPU one; SlabSession checks that."""
    ext = (tree_pos.max(dim=0).values - tree_pos.min(dim=0).values).double()
    area = float(2 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0]))
    if area <= 0:
        area = float(ext.max()) ** 2
    spacing = math.sqrt(area / tree_pos.size(0))
    return factor * spacing * math.sqrt(k / math.pi)


# ------------------------------------------------------------------------------------------------------
# host staging next to the GPU
# ------------------------------------------------------------------------------------------------------
def gpu_local_cpus(device_index: int) -> set[int]:
    """CPUs on the NUMA node the GPU hangs off (NVML's ideal CPU affinity), restricted to the CPUs this process may use.
    Empty when NVML or the topology says nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        return cpus & os.sched_getaffinity(0)
    except Exception:
        return set()


@contextlib.contextmanager
def near_gpu(device_index: int):
    """Run the body on the GPU's own NUMA node.  Pinned host buffers allocated inside land in that node's memory (first
    touch), so their DMA does not cross the socket interconnect when one process per GPU runs on a multi-socket box.  A
    no-op where NVML reports no affinity or the box has one node (the measured boxes: profiles/README.md, e2e note)."""
    cpus = gpu_local_cpus(device_index)
    before = os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
    try:
        yield cpus
    finally:
        if cpus:
            os.sched_setaffinity(0, before)


# ------------------------------------------------------------------------------------------------------
# binding to the CUDA session
# ------------------------------------------------------------------------------------------------------
class _DevView:
    """zero-copy torch view of a raw device pointer (for all-reducing the session's scalar buffers)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": shape, "typestr": typestr, "version": 2}


class SlabSession:
    """One rank's share of a denoising run: owned slab + halo in one CUDA session, halos refreshed between phases."""

    def __init__(self, pos: torch.Tensor, nrm: torch.Tensor, k_feature=16, k_update=8, alphas=(1.0, 0.2, 1.0),
                 strategy=None, halo_width: float | None = None, group=None, tree_pos: torch.Tensor | None = None,
                 transport: str | None = None):
        from . import _lib
        self._lib = _lib
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        tree = pos if tree_pos is None else tree_pos
        hw = halo_width if halo_width is not None else estimate_halo_width(tree, k_feature)
        self.plan = SlabPlan(tree, rank, world, hw)
        self.n_owned, self.n_halo, self.n_total = self.plan.n_owned, self.plan.n_halo, self.plan.n
        ids = self.plan.local_ids
        self.session = _lib.Session(tree[ids].contiguous(), k_feature)
        self.session.set_state(pos[ids].contiguous(), nrm[ids].contiguous())
        self.ex = HaloExchanger(self.plan, group)
        dev = pos.device
        # local index -> session tree position
        perm = self.session.order().long()                       # tree position -> local index
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.numel(), device=dev)
        owned_local = torch.zeros(perm.numel(), dtype=torch.uint8, device=dev)
        owned_local[:self.n_owned] = 1
        self._owned_tree = owned_local[perm].contiguous()
        lib = _lib.load()
        _lib.check(lib.ngpd_session_set_owned(self.session._h, self._owned_tree.data_ptr(), _lib.stream()), "ngpd_session_set_owned")
        self._send_rows = torch.cat([inv[r] for r in self.ex.send_local]).to(torch.int32).contiguous()
        self._recv_rows = torch.cat([inv[r] for r in self.ex.recv_local]).to(torch.int32).contiguous()
        self._send_buf = torch.empty((self._send_rows.numel(), 4), dtype=torch.float32, device=dev)
        self._recv_buf = torch.empty((self._recv_rows.numel(), 4), dtype=torch.float32, device=dev)
        # halo transport: "nccl" = gather kernel + one all-to-all + scatter kernel; "peer" = the gather kernel stores into the
        # peers' receive buffers itself (symmetric memory over NVLink) and a cross-rank barrier replaces the collective
        transport = transport or os.environ.get("NGPD_HALO_TRANSPORT", "nccl")
        assert transport in ("nccl", "peer"), transport
        self.transport = transport if world > 1 else "nccl"
        if self.transport == "peer":
            self.ex.enable_peer_push(dev)
        self._inv = inv
        # global d = 2 * mean 6-NN edge length (Processor.py:120-121)
        s, c = self.session.mean_edge_length_parts(6)
        t = torch.tensor([s, c], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, group=group)
        self.mean_edge_length = float(t[0] / t[1])
        st = strategy if strategy is not None else (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE)
        self.params = _lib.make_params(k_feature, k_update, None, 0.3, 3.0, 0.2, st, alphas, 2.0 * self.mean_edge_length)
        self.world, self.rank = world, rank
        self.exchanges = 0

    # -- halo refresh of session buffer `which` (0 positions, 1 normals, 2 smoothed normals) -----------------
    def _refresh(self, which: int):
        if self.world == 1:
            return
        lib, L = self._lib.load(), self._lib
        h = self.session._h
        if self.transport == "peer":
            ex = self.ex
            b = ex._parity
            ex._parity ^= 1
            L.check(lib.ngpd_session_export_rows_peers(h, which, self._send_rows.data_ptr(), self._send_rows.numel(), ex._seg.data_ptr(),
                                                       ex._peer_base[b].data_ptr(), self.world, L.stream()), "ngpd_session_export_rows_peers")
            ex._hdl.barrier(channel=0)                          # every rank's stores have landed before anybody unpacks
            L.check(lib.ngpd_session_import_rows(h, which, self._recv_rows.data_ptr(), self._recv_rows.numel(), ex._sym[b].data_ptr(), L.stream()),
                    "ngpd_session_import_rows")
            self.exchanges += 1
            return
        L.check(lib.ngpd_session_export_rows(h, which, self._send_rows.data_ptr(), self._send_rows.numel(), self._send_buf.data_ptr(), L.stream()),
                "ngpd_session_export_rows")
        self.ex.exchange(self._send_buf, self._recv_buf)
        L.check(lib.ngpd_session_import_rows(h, which, self._recv_rows.data_ptr(), self._recv_rows.numel(), self._recv_buf.data_ptr(), L.stream()),
                "ngpd_session_import_rows")
        self.exchanges += 1

    def _scalar_view(self, which, shape, typestr):
        ptr = self._lib.load().ngpd_session_buffer(self.session._h, which)
        return torch.as_tensor(_DevView(ptr, shape, typestr), device=self._send_buf.device)

    def _views(self):
        # the flat-step accumulators live at fixed addresses for the session's life: wrap them once
        if not hasattr(self, "_scalar_views"):
            self._scalar_views = (self._scalar_view(3, (4,), "<f8"), self._scalar_view(4, (4,), "<f4")[3:4])
        return self._scalar_views

    def step(self):
        import ctypes
        lib, L, p = self._lib.load(), self._lib, self.params
        h, st = self.session._h, L.stream
        self.launches = 0
        ref = ctypes.byref(p)
        L.check(lib.ngpd_session_phase_features(h, ref, 0, st()), "phase_features 0")
        self._refresh(2)                                        # neighbours' smoothed normals
        L.check(lib.ngpd_session_phase_features(h, ref, 1, st()), "phase_features 1")
        for key in range(3):
            kind = p.strategy[key]
            if kind < 0:
                continue
            if kind == L.STEP_FLAT:
                L.check(lib.ngpd_session_phase_flat_scalars(h, ref, key, 0, st()), "flat scalars 0")
                if self.world > 1:
                    dist.all_reduce(self._views()[0], group=self.group)
                L.check(lib.ngpd_session_phase_flat_scalars(h, ref, key, 1, st()), "flat scalars 1")
                if self.world > 1:
                    dist.all_reduce(self._views()[1], op=dist.ReduceOp.MAX, group=self.group)
            L.check(lib.ngpd_session_phase_update(h, ref, key, st()), "phase_update")
            self._refresh(0)                                    # the class' new positions, before the next class reads them
        L.check(lib.ngpd_session_phase_commit_normals(h), "commit normals")

    def owned_state(self):
        """(original ids, positions, normals, labels) of the owned points"""
        pos, nrm, lab = self.session.get_state(True)
        k = self.n_owned
        return self.plan.owned, pos[:k], nrm[:k], lab[:k]

    def gather_full(self):
        """positions / normals / labels of the whole cloud in original order on every rank (tests, small clouds)"""
        ids, pos, nrm, lab = self.owned_state()
        n, dev = self.n_total, pos.device
        full_p = torch.zeros((n, 3), device=dev); full_n = torch.zeros((n, 3), device=dev)
        full_l = torch.zeros(n, dtype=torch.int32, device=dev)
        full_p[ids] = pos; full_n[ids] = nrm; full_l[ids] = lab.int()
        if self.world > 1:
            for t in (full_p, full_n, full_l):
                dist.all_reduce(t, group=self.group)
        return full_p, full_n, full_l.to(torch.uint8)

    # -- end-to-end step with HOST buffers of the OWNED rows (what bench.py's e2e times on more than one GPU) -----------
    def step_host(self, pos_host: torch.Tensor, nrm_host: torch.Tensor, pos_out: torch.Tensor, nrm_out: torch.Tensor,
                  lab_out: torch.Tensor):
        """One iteration whose inputs and outputs live in pinned host memory: this rank's owned positions / normals
        ([n_owned,3] fp32, slab order = `plan.owned`) go host -> device, halo copies are refreshed from their owners,
        the iteration runs with its halo exchanges, and the owned positions / normals / labels come back."""
        lib, L = self._lib.load(), self._lib
        h, dev, k = self.session._h, self._send_buf.device, self.n_owned
        if not hasattr(self, "_owned_rows"):
            self._owned_rows = self._inv[:k].to(torch.int32).contiguous()       # tree positions of the owned rows
            self._io4 = torch.zeros((k, 4), dtype=torch.float32, device=dev)
        rows, io4 = self._owned_rows, self._io4
        for which, src in ((0, pos_host), (1, nrm_host)):
            io4[:, :3].copy_(src, non_blocking=True)
            L.check(lib.ngpd_session_import_rows(h, which, rows.data_ptr(), k, io4.data_ptr(), L.stream()), "ngpd_session_import_rows")
            self._refresh(which)
        self.step()
        for which, dst in ((0, pos_out), (1, nrm_out)):
            L.check(lib.ngpd_session_export_rows(h, which, rows.data_ptr(), k, io4.data_ptr(), L.stream()), "ngpd_session_export_rows")
            dst.copy_(io4[:, :3], non_blocking=True)
        labels = self._scalar_view(5, (self.n_owned + self.n_halo,), "|u1")
        lab_out.copy_(labels[rows.long()], non_blocking=True)
        torch.cuda.current_stream().synchronize()
