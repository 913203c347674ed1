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



# ------------------------------------------------------------------------------------------------------
# planning from SHARDED input: no rank ever holds more than its slab + halo (+ the shard it was handed)
# ------------------------------------------------------------------------------------------------------
def _route(tensors, dest: torch.Tensor, world: int, group):
    """Send row i of every tensor to rank dest[i] (one variable-split all-to-all per tensor).  Returns (received tensors,
    rows received from each rank); rows arrive grouped by source rank, in the source's order."""
    order = torch.argsort(dest, stable=True)
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = [int(v) for v in send_counts.tolist()], [int(v) for v in recv_counts.tolist()]
    out = []
    for t in tensors:
        src = t[order].contiguous()
        dst = torch.empty((sum(rc),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(dst, src, output_split_sizes=rc, input_split_sizes=sc, group=group)
        out.append(dst)
    return out, rc


def _all_gather_var(t: torch.Tensor, world: int, group):
    """all-gather of 1-D tensors of different lengths -> list of tensors"""
    size = torch.tensor([t.numel()], dtype=torch.long, device=t.device)
    sizes = [torch.empty_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(v.item()) for v in sizes]
    pad = torch.zeros(max(max(sizes), 1), dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:m] for b, m in zip(bufs, sizes)]


def sample_splitters(keys: torch.Tensor, world: int, group, samples: int = 1 << 16, fractions=None) -> torch.Tensor:
    """world - 1 Morton keys that cut the union of every rank's `keys` into world slabs of nearly equal size (or of the given
    `fractions` of the points): every rank contributes `samples` evenly spaced quantiles of its own keys, weighted by the rows
    each stands for, and the cuts are read off the merged sample.  Imbalance <= world^2 / samples of a slab."""
    m = keys.numel()
    srt = torch.sort(keys).values
    take = min(samples, m)
    if take > 0:
        pick = ((torch.arange(take, device=keys.device, dtype=torch.float64) + 0.5) * (m / take)).long().clamp_(0, m - 1)
        smp, wgt = srt[pick], torch.full((take,), m / take, dtype=torch.float64, device=keys.device)
    else:
        smp, wgt = srt[:0], torch.zeros(0, dtype=torch.float64, device=keys.device)
    all_s = torch.cat(_all_gather_var(smp, world, group))
    all_w = torch.cat(_all_gather_var(wgt, world, group))
    o = torch.argsort(all_s, stable=True)
    all_s, cum = all_s[o], torch.cumsum(all_w[o], 0)
    total = float(cum[-1]) if cum.numel() else 0.0
    if fractions is None:
        fractions = [1.0 / world] * world
    assert len(fractions) == world and min(fractions) > 0
    norm, run, cuts = float(sum(fractions)), 0.0, []
    for f in list(fractions)[:-1]:
        run += float(f) / norm
        cuts.append(total * run)
    targets = torch.tensor(cuts, dtype=torch.float64, device=keys.device)
    at = torch.searchsorted(cum, targets).clamp_(max=max(all_s.numel() - 1, 0))
    return all_s[at] if all_s.numel() else torch.zeros(world - 1, dtype=torch.long, device=keys.device)


def rebalance_fractions(times, fractions, damping: float = 1.0):
    """Cost-aware slab sizes.  `times[r]` = measured kernel time of the slab that held `fractions[r]` of the points (slabs are
    contiguous in Morton order): taking the cost per point as constant inside each old slab, returns the fractions whose
    slabs all cost the same.  (The cost of a row varies by region -- rows that leave the re-ranking tier of the k-NN cost
    ~20 x, crease rows take the LAPACK-order path and the listed-rows updates -- and the phases of an iteration are separated
    by cross-rank rounds, so the slowest slab sets the pace.)  damping < 1 moves only part of the way."""
    world = len(times)
    tot_f = float(sum(fractions))
    f = [float(x) / tot_f for x in fractions]
    t = [max(float(x), 1e-12) for x in times]
    total, target = sum(t), sum(t) / world
    # piecewise-linear cumulative cost over the point coordinate x in [0, 1]; invert it at j * total / world
    edges_x, edges_c = [0.0], [0.0]
    for fr, tr in zip(f, t):
        edges_x.append(edges_x[-1] + fr)
        edges_c.append(edges_c[-1] + tr)
    cuts = [0.0]
    seg = 0
    for j in range(1, world):
        c = j * target
        while seg < world - 1 and edges_c[seg + 1] < c:
            seg += 1
        a = (c - edges_c[seg]) / (edges_c[seg + 1] - edges_c[seg])
        cuts.append(edges_x[seg] + a * (edges_x[seg + 1] - edges_x[seg]))
    cuts.append(1.0)
    new = [cuts[j + 1] - cuts[j] for j in range(world)]
    out = [(1.0 - damping) * o + damping * nw for o, nw in zip(f, new)]
    s_ = sum(out)
    return [x / s_ for x in out]


class ShardedSlabPlan:
    """The plan of SlabPlan, built from input that is already spread over the ranks: every rank passes the construction-time
    positions of ITS shard (any subset; `gids` are the points' ids in the whole cloud) plus per-point payload tensors.
    Points are routed to the slab that owns their Morton key, halo copies are pushed by their owners.  Communication:
    a bounding-box all-reduce, one key sample all-gather, and variable-split all-to-alls; memory per rank: shard + slab + halo.
    Fields as SlabPlan (`owned`, `halo`, `local_ids` hold global ids) plus the routed data: `tree_local` [n_owned + n_halo, 3]
    and `payload_local`, and the halo wiring itself (`send_local`, `recv_counts`), so HaloExchanger needs no request round."""

    def __init__(self, tree_pos: torch.Tensor, gids: torch.Tensor, halo_width: float, group=None, payload=(), fractions=None):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = tree_pos.device
        self.rank, self.world, self.halo_width = rank, world, float(halo_width)
        self.fractions = list(fractions) if fractions is not None else [1.0 / world] * world
        lo = tree_pos.min(dim=0).values if tree_pos.numel() else torch.full((3,), float("inf"), device=dev)
        hi = tree_pos.max(dim=0).values if tree_pos.numel() else torch.full((3,), float("-inf"), device=dev)
        cnt = torch.tensor([tree_pos.size(0)], dtype=torch.long, device=dev)
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(cnt, group=group)
        self.n = int(cnt.item())
        self.lo, self.hi = lo, hi
        keys = morton_keys(tree_pos, lo, hi)
        data = [gids.long(), tree_pos] + list(payload)
        if world > 1:
            self.splitters = sample_splitters(keys, world, group, fractions=fractions)
            dest = torch.bucketize(keys, self.splitters, right=True)
            del keys
            data, _ = _route(data, dest, world, group)
            del dest
        o = torch.argsort(data[0])                                  # deterministic local order whatever the sharding was
        data = [t[o] for t in data]
        self.owned = data[0]
        self.n_owned = self.owned.numel()
        own_tree = data[1]
        # halo: every owned point whose coarse cell (edge = halo width) lies in the 27-neighbourhood of a cell another rank occupies
        w = self.halo_width
        cell = ((own_tree.double() - lo.double()) / w).floor_().long() + 1
        dims = ((hi.double() - lo.double()) / w).floor_().long() + 3
        lin = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
        self.send_local = [torch.empty(0, dtype=torch.long, device=dev) for _ in range(world)]
        halo_data = [t[:0] for t in data]
        self.recv_counts = [0] * world
        if world > 1:
            mine = torch.unique(lin)
            offs = torch.tensor([(dz * int(dims[1]) + dy) * int(dims[0]) + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)],
                                device=dev)
            near = torch.unique((mine[:, None] + offs[None, :]).reshape(-1))
            nears = _all_gather_var(near, world, group)
            rows, dests = [], []
            for q in range(world):
                if q == rank:
                    continue
                self.send_local[q] = torch.isin(lin, nears[q]).nonzero().flatten()
                rows.append(self.send_local[q])
                dests.append(torch.full_like(self.send_local[q], q))
            rows, dests = torch.cat(rows), torch.cat(dests)
            halo_data, self.recv_counts = _route([t[rows] for t in data], dests, world, group)
        self.halo = halo_data[0]
        self.n_halo = self.halo.numel()
        self.halo_owner = torch.repeat_interleave(torch.arange(world, device=dev), torch.tensor(self.recv_counts, device=dev))
        self.local_ids = torch.cat([self.owned, self.halo])
        self.tree_local = torch.cat([own_tree, halo_data[1]])
        self.payload_local = [torch.cat([a, b]) for a, b in zip(data[2:], halo_data[2:])]


class HaloExchanger:
    """Fixed send/recv wiring between slabs.  `send_local[q]` = local indices (of owned points) whose values go to
    peer q; `recv_local[q]` = local indices (of halo points) that peer q fills, in matching order."""

    def __init__(self, plan: SlabPlan, group=None):
        self.plan, self.group = plan, group
        world, rank, dev = plan.world, plan.rank, plan.local_ids.device
        if hasattr(plan, "send_local"):
            # ShardedSlabPlan: the owners chose the halo rows themselves, nothing to ask for
            self.send_local = plan.send_local
            self.send_counts = [int(r.numel()) for r in plan.send_local]
            self.recv_counts = list(plan.recv_counts)
            self.recv_local, off = [], plan.n_owned
            for q in range(world):
                self.recv_local.append(torch.arange(off, off + self.recv_counts[q], device=dev))
                off += self.recv_counts[q]
            return
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

    # -- halo push over peer memory (SlabSession(transport="peer"), the default on GPUs) ---------------------------------------
    def peer_layout(self, device: torch.device):
        """(cap, first_row, seg) of the peer-memory push, from one all-gather of the receive counts (peer_push_layout)."""
        world, rank = self.plan.world, self.plan.rank
        mine = torch.tensor(self.recv_counts, dtype=torch.long, device=device)
        table = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(table, mine, group=self.group)
        recv = [[int(v) for v in t.tolist()] for t in table]              # recv[q][s] = rows rank q receives from rank s
        cap, first_row, seg = peer_push_layout(recv, rank)
        assert seg[1:] == [sum(self.send_counts[:q + 1]) for q in range(world)]
        return cap, first_row, seg

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
    return halo_width_from_box(tree_pos.min(dim=0).values, tree_pos.max(dim=0).values, tree_pos.size(0), k, factor)


def halo_width_from_box(lo: torch.Tensor, hi: torch.Tensor, n: int, k: int, factor: float = 6.0) -> float:
    ext = (hi - lo).double()
    area = float(2 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0]))
    if area <= 0:
        area = float(ext.max()) ** 2
    spacing = math.sqrt(area / max(n, 1))
    return factor * spacing * math.sqrt(k / math.pi)


def estimate_halo_width_sharded(tree_pos: torch.Tensor, k: int, group=None, factor: float = 6.0) -> float:
    """estimate_halo_width when every rank holds only a shard of the cloud"""
    dev = tree_pos.device
    lo = tree_pos.min(dim=0).values if tree_pos.numel() else torch.full((3,), float("inf"), device=dev)       # (an empty shard is a valid shard)
    hi = tree_pos.max(dim=0).values if tree_pos.numel() else torch.full((3,), float("-inf"), device=dev)
    cnt = torch.tensor([tree_pos.size(0)], dtype=torch.long, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(cnt, group=group)
    return halo_width_from_box(lo, hi, int(cnt.item()), k, factor)


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

    def __init__(self, pos: torch.Tensor, nrm: torch.Tensor | None, k_feature=16, k_update=8, alphas=(1.0, 0.2, 1.0),
                 strategy=None, halo_width: float | None = None, group=None, tree_pos: torch.Tensor | None = None,
                 transport: str | None = None, shard_ids: torch.Tensor | None = None, mean_edge_length: float | None = None,
                 flags: int = 0, clamp_radius: float = 0.0, fractions=None):
        """Replicated input (shard_ids None): every rank passes the WHOLE cloud and keeps its slab (small clouds, tests).
        Sharded input: every rank passes any part of the cloud with the points' global ids in `shard_ids`; rows are routed
        to their slabs (ShardedSlabPlan) and no rank holds more than shard + slab + halo.  nrm None: normals are set later
        (set_owned_normals / pca_normals).  fractions (sharded input): share of the points per slab, e.g. from
        rebalance_fractions() after a trial run; default equal."""
        from . import _lib
        self._lib = _lib
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        tree = pos if tree_pos is None else tree_pos
        if shard_ids is None:
            hw = halo_width if halo_width is not None else estimate_halo_width(tree, k_feature)
            self.plan = SlabPlan(tree, rank, world, hw)
            ids = self.plan.local_ids
            tree_l, pos_l, nrm_l = tree[ids].contiguous(), pos[ids].contiguous(), (nrm[ids].contiguous() if nrm is not None else None)
        else:
            hw = halo_width if halo_width is not None else estimate_halo_width_sharded(tree, k_feature, group)
            payload = ([pos] if tree_pos is not None else []) + ([nrm] if nrm is not None else [])
            self.plan = ShardedSlabPlan(tree, shard_ids, hw, group, payload, fractions)
            tree_l = self.plan.tree_local.contiguous()
            rest = list(self.plan.payload_local)
            pos_l = rest.pop(0).contiguous() if tree_pos is not None else tree_l
            nrm_l = rest.pop(0).contiguous() if nrm is not None else None
        self.n_owned, self.n_halo, self.n_total = self.plan.n_owned, self.plan.n_halo, self.plan.n
        self.tree_local = tree_l
        self.session = _lib.Session(tree_l, k_feature)
        self.session.set_state(pos_l, nrm_l)
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
        # halo transport.  "peer" (default): the whole step is driven from C (ngpd_session_step_slab); halo rows are stored
        # straight into the peers' receive buffers (symmetric memory over NVLink / NVSwitch), signalled with release stores and
        # awaited with acquire loads, and flat_step's two cloud-wide scalars go through the same block -- no NCCL call, no host
        # round trip inside an iteration.  "nccl": gather kernel + one all-to-all + scatter kernel per refresh and two
        # all-reduces, phase by phase from Python (the round-1 path; also what the gloo tests exercise on the CPU side).
        transport = transport or os.environ.get("NGPD_HALO_TRANSPORT", "peer")
        assert transport in ("nccl", "peer"), transport
        self.transport = transport if world > 1 else "nccl"
        if self.transport == "peer":
            self._wire_peer(dev)
        self._inv = inv
        self.world, self.rank = world, rank
        self.exchanges = 0
        # global d = 2 * mean 6-NN edge length (Processor.py:120-121)
        if mean_edge_length is None:
            s, c = self.session.mean_edge_length_parts(6)
            t = torch.tensor([s, c], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, group=group)
            mean_edge_length = float(t[0] / t[1])
        self.mean_edge_length = mean_edge_length
        st = strategy if strategy is not None else (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE)
        self.params = _lib.make_params(k_feature, k_update, None, 0.3, 3.0, 0.2, st, alphas, 2.0 * self.mean_edge_length, flags, clamp_radius)

    def set_owned_normals(self, nrm_owned: torch.Tensor):
        """normals of the owned rows (slab order = plan.owned); the halo copies are fetched from their owners"""
        full = torch.zeros((self.n_owned + self.n_halo, 3), dtype=torch.float32, device=nrm_owned.device)
        full[:self.n_owned] = nrm_owned
        self.session.set_state(None, full)
        self._refresh(1)

    def pca_normals(self, k: int = 12, orient_like=None) -> torch.Tensor:
        """GraphBuilder.getKNNEdgeIndex(k) + setPVTNormals on the slab (GraphBuilder.py:60-63, 95-111): k-NN graph without
        self over the construction-time positions, PCA normal per owned row, optionally flipped to agree with `orient_like`
        (owned rows); sets them as the session's normals and returns them."""
        L = self._lib
        tree = self.tree_local
        grid = L.Grid(tree, k)
        table = grid.knn(tree[:self.n_owned].contiguous(), k, L.KNN_SKIP_SELF)
        nrm = torch.empty((self.n_owned, 3), dtype=torch.float32, device=tree.device)
        with torch.cuda.device(tree.device):
            L.check(L.load().ngpd_pca_normals(tree.data_ptr(), table.data_ptr(), None, self.n_owned, k, nrm.data_ptr(), None, None, L.stream()),
                    "ngpd_pca_normals")
        if isinstance(orient_like, str):
            assert orient_like == "current"                     # the normals the session holds now (e.g. analytic ones routed with the shards)
            orient_like = self.owned_state()[2]
        if orient_like is not None:
            flip = (nrm * orient_like).sum(1) < 0
            nrm[flip] *= -1
        del grid, table
        self.set_owned_normals(nrm)
        return nrm

    def checksum(self) -> list[int]:
        """digest of the whole cloud's state (every rank gets it): per-slab digests add up modulo 2^64"""
        part = self.session.checksum(self.plan.local_ids)
        t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in part], dtype=torch.int64, device=self._send_buf.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return [int(v) & ((1 << 64) - 1) for v in t.tolist()]

    def _wire_peer(self, dev):
        """One block of symmetric memory per rank (receive buffers of the halo rows, scalar slots, round flags: include/ngpd.h,
        ngpd_slab_wiring_t), mapped into every peer; the addresses go to the session once."""
        import ctypes
        import torch.distributed._symmetric_memory as symm
        L, lib = self._lib, self._lib.load()
        world, rank = self.plan.world, self.plan.rank
        group = self.group if self.group is not None else dist.group.WORLD
        cap, first_row, seg = self.ex.peer_layout(dev)
        nbytes = int(lib.ngpd_slab_symm_bytes(world, cap))
        assert nbytes > 0
        self._sym = symm.empty(((nbytes + 15) // 16 * 16,), dtype=torch.uint8, device=dev)
        self._sym.zero_()
        self._hdl = symm.rendezvous(self._sym, group)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                              # everybody's flags are zero before anybody signals
        ptrs = [int(a) for a in self._hdl.buffer_ptrs]
        w = L.SlabWiring()
        w.world, w.rank, w.n_send, w.n_recv, w.cap = world, rank, self._send_rows.numel(), self._recv_rows.numel(), cap
        w.send_rows, w.recv_rows = self._send_rows.data_ptr(), self._recv_rows.data_ptr()
        self._wire_keep = ((ctypes.c_int64 * (world + 1))(*seg), (ctypes.c_int64 * world)(*first_row), (ctypes.c_uint64 * world)(*ptrs))
        w.send_seg_host, w.first_row_host, w.symm_base_host = self._wire_keep
        with torch.cuda.device(dev):
            L.check(lib.ngpd_session_set_slab(self.session._h, ctypes.byref(w), L.stream()), "ngpd_session_set_slab")

    # -- halo refresh of session buffer `which` (0 positions, 1 normals, 2 smoothed normals) -----------------
    def _refresh(self, which: int):
        if self.world == 1:
            return
        lib, L = self._lib.load(), self._lib
        h = self.session._h
        if self.transport == "peer":
            L.check(lib.ngpd_session_slab_refresh(h, which, L.stream()), "ngpd_session_slab_refresh")
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
            # buffer 3: flat_step's centre sums as int64 fixed point (exact: any summation order gives the same bits)
            self._scalar_views = (self._scalar_view(3, (4,), "<i8"), self._scalar_view(4, (4,), "<f4")[3:4])
        return self._scalar_views

    def step(self):
        import ctypes
        lib, L, p = self._lib.load(), self._lib, self.params
        h, st = self.session._h, L.stream
        self.launches = 0
        ref = ctypes.byref(p)
        if self.transport == "peer":
            L.check(lib.ngpd_session_step_slab(h, ref, st()), "ngpd_session_step_slab")
            return
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

    def rank_costs(self, steps: int = 3, skip: int = 2):
        """Runs `skip` + `steps` iterations and returns every rank's kernel time per step (ms) over the last `steps`, the
        cross-rank rounds excluded -- what rebalance_fractions() wants.  The first iterations are skipped because their search
        has no stored candidates and costs several times a steady-state one, with another distribution over the cloud.
        The state advances; use a throw-away session."""
        s = self.session
        for _ in range(skip):
            self.step()
        s.set_profiling(True)
        s.get_profile()
        for _ in range(steps):
            self.step()
        prof = s.get_profile()
        s.set_profiling(False)
        mine = torch.tensor([sum(v[0] for k, v in prof.items() if k != "halo") / steps], dtype=torch.float64, device=self._send_buf.device)
        table = [torch.empty_like(mine) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(table, mine, group=self.group)
        else:
            table = [mine]
        return [float(t.item()) for t in table]

    def verify_halo(self) -> float:
        """The slab searches equal the whole cloud's iff every owned row's (k-th neighbour distance + displacement from its
        tree position) stayed below the halo width (the session tracks the maximum, ngpd_session_halo_need).  Returns the
        largest value any rank has seen; raises on every rank when it reached the width."""
        need = torch.tensor([self.session.halo_need()], dtype=torch.float64, device=self._send_buf.device)
        if self.world > 1:
            dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        need = float(need.item())
        if not need < self.plan.halo_width:
            raise RuntimeError(f"halo too narrow: a query needed points up to {need:.6g} from its tree position, the halo is "
                               f"{self.plan.halo_width:.6g} wide -- rebuild the SlabSession with halo_width > {need:.6g}")
        return need

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


# ------------------------------------------------------------------------------------------------------
# sharded metrics and k-NN (BASELINE configs[4] beyond one GPU)
# ------------------------------------------------------------------------------------------------------
def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    size = torch.tensor([t.size(0)], dtype=torch.long, device=t.device)
    sizes = [torch.empty_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(v.item()) for v in sizes]
    pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.size(0)] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:m] for b, m in zip(bufs, sizes)])


def sharded_chamfer(a_shard: torch.Tensor, b_shard: torch.Tensor, group=None) -> dict:
    """TorchUtils.ChamferDistance(A, B).mean() and friends (Utils.py:253-295) for clouds spread over the ranks: every rank
    queries ITS shard of B against a replica of A and its shard of A against a replica of B (a replicated target
    costs 12 B/point -- 12 GB at 10^9 points -- which a B200 holds), with the block reduction fused into the nearest-
    neighbour kernel (ngpd_nn_sqdist_reduce: no per-point output), then one all-reduce of 8 doubles.  Returns python floats."""
    from . import _lib
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    a_full = _all_gather_rows(a_shard, group) if multi else a_shard
    ga = _lib.Grid(a_full, 4)
    del a_full
    acc_b = ga.nn_reduce(b_shard)                               # every B point -> nearest A point (first half of the reference's vector)
    del ga
    b_full = _all_gather_rows(b_shard, group) if multi else b_shard
    gb = _lib.Grid(b_full, 4)
    del b_full
    acc_a = gb.nn_reduce(a_shard)
    del gb
    sums = torch.stack([acc_b[0], acc_b[1], acc_b[3], acc_a[0], acc_a[1], acc_a[3]])
    mx = torch.stack([acc_b[2], acc_a[2]])
    if multi:
        dist.all_reduce(sums, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    s = [float(v) for v in sums.tolist()]
    return {"chamfer": (s[0] + s[3]) / (s[2] + s[5]), "single_chamfer": s[0] / s[2], "hausdorff": math.sqrt(float(mx.max())),
            "mean_distance": s[1] / s[2], "rows": int(s[2] + s[5])}


class ShardedKnn:
    """k nearest neighbours of a cloud that is spread over the ranks (Selector.getKNNSelection beyond one GPU): Morton slabs
    from the shards (ShardedSlabPlan), one index per rank over owned + halo rows, every rank answers its owned rows.  Rows come
    back as GLOBAL ids.  `check()` proves the rows equal the whole cloud's (k-th distance below the halo width)."""

    def __init__(self, shard_pos: torch.Tensor, shard_ids: torch.Tensor, k_max: int, group=None, halo_width: float | None = None):
        from . import _lib
        self._lib, self.group = _lib, group
        hw = halo_width if halo_width is not None else estimate_halo_width_sharded(shard_pos, k_max, group)
        self.plan = ShardedSlabPlan(shard_pos, shard_ids, hw, group)
        self.grid = _lib.Grid(self.plan.tree_local, k_max)
        self.n_owned = self.plan.n_owned
        self.queries = self.plan.tree_local[:self.n_owned].contiguous()

    def knn(self, k: int, flags: int = 0, global_ids: bool = True, with_d2: bool = False):
        out = self.grid.knn(self.queries, k, flags, with_d2=True)
        idx, d2 = out
        self._last_dk = d2[:, k - 1].max() if idx.numel() else torch.zeros((), device=idx.device)
        if global_ids:
            idx = self.plan.local_ids[idx.long()]
        return (idx, d2) if with_d2 else idx

    def check(self) -> float:
        need = self._last_dk.double().sqrt().reshape(1)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        need = float(need.item())
        if not need < self.plan.halo_width:
            raise RuntimeError(f"halo too narrow: k-th neighbour at {need:.6g}, halo width {self.plan.halo_width:.6g}")
        return need
