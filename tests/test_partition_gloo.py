"""Host-side logic of the multi-GPU path (Morton slabs, halo membership, request wiring, halo refresh, scalar
all-reduces) exercised with world_size = 2 over gloo on CPU.  The per-point kernels are not involved: values are
synthetic functions of the point id, and neighbourhoods come from the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cloud(n=12000, seed=4):
    rng = np.random.default_rng(seed)
    u, v = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    p = np.stack([(0.6 + 0.25 * np.cos(v)) * np.cos(u), (0.6 + 0.25 * np.cos(v)) * np.sin(u), 0.25 * np.sin(v)], 1)
    return (p + rng.normal(0, 0.002, p.shape)).astype(np.float32)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ngpd_oracle as O
        from ngpd_b200 import partition
        pos_np = _cloud()
        pos = torch.from_numpy(pos_np)
        n = len(pos_np)
        k = 16
        hw = partition.estimate_halo_width(pos, k, factor=3.0)
        plan = partition.SlabPlan(pos, rank, world, hw)
        ex = partition.HaloExchanger(plan)

        # 1. slabs partition the cloud, equal sizes
        owned = [None] * world
        dist.all_gather_object(owned, plan.owned.tolist())
        flat = np.concatenate([np.asarray(o) for o in owned])
        assert len(flat) == n and len(np.unique(flat)) == n
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        assert bool((plan.slab_of[plan.owned] == rank).all()) and bool((plan.slab_of[plan.halo] != rank).all())

        # 2. the halo holds every foreign point a moved owned query can reach: move each point by up to 2 spacings and
        #    check its exact 16-NN (oracle, full cloud) lies inside owned + halo
        rng = np.random.default_rng(rank)
        spacing = hw / (3.0 * np.sqrt(k / np.pi))
        moved = pos_np[plan.owned.numpy()] + rng.normal(0, 0.5 * spacing, (plan.n_owned, 3)).astype(np.float32)
        nn = O.knn_bruteforce(pos_np, moved, k)
        local = np.zeros(n, dtype=bool)
        local[plan.local_ids.numpy()] = True
        assert local[nn].all(), "halo too thin"
        assert plan.n_halo < 0.5 * n                                        # and it is a boundary layer, not the whole cloud

        # 3. halo refresh: values that are a function of the original id arrive in the right rows
        def values(ids, phase):
            return torch.stack([ids.float() * (phase + 1), ids.float() + 0.25, -ids.float(), torch.full_like(ids, phase).float()], 1)

        state = torch.zeros((plan.local_ids.numel(), 4))
        for phase in range(3):
            state[:plan.n_owned] = values(plan.owned, phase)
            send = torch.cat([state[r] for r in ex.send_local])
            recv = torch.empty((sum(ex.recv_counts), 4))
            ex.exchange(send, recv)
            state[torch.cat(ex.recv_local)] = recv
            assert torch.equal(state, values(plan.local_ids, phase)), f"phase {phase}"
        sb, rb = ex.bytes_per_exchange()
        assert sb == sum(ex.send_counts) * 16 and rb == plan.n_halo * 16

        # 4. flat_step's cloud-wide scalars from per-rank partial sums (sum, then max): same as the global ones
        vj = pos[plan.owned].double()
        acc = torch.cat([vj.sum(0), torch.tensor([float(plan.n_owned)], dtype=torch.float64)])
        dist.all_reduce(acc)
        centre = (acc[:3] / acc[3]).float()
        delta = (pos[plan.owned] - centre).norm(dim=1).max()
        dist.all_reduce(delta, op=dist.ReduceOp.MAX)
        gc = pos.double().mean(0).float()
        assert torch.allclose(centre, gc, atol=1e-6) and abs(float(delta) - float((pos - gc).norm(dim=1).max())) < 1e-6

        # 5. a distributed neighbourhood average over two phases equals the single-process result
        nbr_full = O.knn_bruteforce(pos_np, pos_np, 8)
        val = np.sin(np.arange(n, dtype=np.float64))
        ref = val.copy()
        for _ in range(2):
            ref = ref[nbr_full].mean(1)
        look = np.full(n, -1); look[plan.local_ids.numpy()] = np.arange(plan.local_ids.numel())
        nbr_local = look[nbr_full[plan.owned.numpy()]]
        assert (nbr_local >= 0).all()
        st = torch.zeros((plan.local_ids.numel(), 4), dtype=torch.float64)
        st[:, 0] = torch.from_numpy(val[plan.local_ids.numpy()])
        for _ in range(2):
            new = st[:, 0].numpy()[nbr_local].mean(1)
            st[:plan.n_owned, 0] = torch.from_numpy(new)
            send = torch.cat([st[r] for r in ex.send_local]); recv = torch.empty((sum(ex.recv_counts), 4), dtype=torch.float64)
            ex.exchange(send, recv)
            st[torch.cat(ex.recv_local)] = recv
        assert np.allclose(st[:plan.n_owned, 0].numpy(), ref[plan.owned.numpy()], rtol=0, atol=1e-15)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_slab_plan_and_halo_exchange_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}


def test_slab_plan_single_rank_has_no_halo():
    sys.path.insert(0, ROOT)
    from ngpd_b200 import partition
    pos = torch.from_numpy(_cloud(2000))
    plan = partition.SlabPlan(pos, 0, 1, partition.estimate_halo_width(pos, 16))
    assert plan.n_owned == 2000 and plan.n_halo == 0
    assert torch.equal(torch.sort(plan.owned).values, torch.arange(2000))


def test_morton_keys_order_is_spatial():
    sys.path.insert(0, ROOT)
    from ngpd_b200 import partition
    pos = torch.rand(20000, 3)
    keys = partition.morton_keys(pos, pos.min(0).values, pos.max(0).values)
    # the leading bit of the code is the leading bit of the quantised z coordinate
    zq = ((pos[:, 2] - pos[:, 2].min()) / (pos.max(0).values - pos.min(0).values).max() * (2 ** 21 - 1)).floor().long()
    assert torch.equal(keys >> 62, zq >> 20)
    # points sharing a long key prefix are close: consecutive points along the curve are near each other on average
    order = torch.argsort(keys)
    step = (pos[order[1:]] - pos[order[:-1]]).norm(dim=1).mean()
    rand = (pos[1:] - pos[:-1]).norm(dim=1).mean()
    assert float(step) < 0.15 * float(rand)


def test_peer_push_layout_emulated():
    """Addressing of the peer-memory halo push (partition.peer_push_layout), emulated on the host for 4 ranks with ragged
    counts: every rank stores its packed send rows at first_row[q] + (i - seg[q]) of peer q's buffer; afterwards each
    buffer must hold, grouped by source rank in ascending order, exactly what an all-to-all would have delivered."""
    from ngpd_b200.partition import peer_push_layout
    rng = np.random.default_rng(3)
    world = 4
    recv = rng.integers(0, 9, size=(world, world))
    np.fill_diagonal(recv, 0)
    recv[2, :] = 0                                                       # a rank that receives nothing
    recv = recv.tolist()
    send_rows = {(s, q): [(s, q, j) for j in range(recv[q][s])] for s in range(world) for q in range(world)}
    caps, buffers = set(), None
    for s in range(world):
        cap, first_row, seg = peer_push_layout(recv, s)
        caps.add(cap)
        if buffers is None:
            buffers = [[None] * cap for _ in range(world)]
        packed = [r for q in range(world) for r in send_rows[(s, q)]]    # this rank's send list, grouped by destination
        assert seg[-1] == len(packed)
        for i, row in enumerate(packed):
            q = max(p for p in range(world) if seg[p] <= i and seg[p] < seg[p + 1] and i < seg[p + 1])
            assert buffers[q][first_row[q] + i - seg[q]] is None          # nobody else writes this slot
            buffers[q][first_row[q] + i - seg[q]] = row
    assert len(caps) == 1                                                # symmetric: one size on every rank
    for q in range(world):
        want = [r for s in range(world) for r in send_rows[(s, q)]]
        assert buffers[q][:len(want)] == want
        assert all(v is None for v in buffers[q][len(want):])


def _sharded_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ngpd_oracle as O
        from ngpd_b200 import partition
        pos_np = _cloud(9000, seed=11)
        n = len(pos_np)
        k = 16
        # an arbitrary, uneven sharding of the input: rank 0 holds a third of the cloud (interleaved ids), rank 1 the rest
        ids = np.arange(n)
        mine = ids[ids % 3 == 0] if rank == 0 else ids[ids % 3 != 0]
        shard = torch.from_numpy(pos_np[mine])
        gid = torch.from_numpy(mine)
        hw = partition.estimate_halo_width_sharded(shard, k, factor=3.0)
        assert abs(hw - partition.estimate_halo_width(torch.from_numpy(pos_np), k, factor=3.0)) < 1e-12
        payload = torch.stack([gid.float() * 0.5, -gid.float()], 1)
        plan = partition.ShardedSlabPlan(shard, gid, hw, None, [payload])
        ex = partition.HaloExchanger(plan)
        assert plan.n == n

        # 1. the slabs partition the cloud into nearly equal parts, rows and payload arrive with their ids
        owned = [None] * world
        dist.all_gather_object(owned, plan.owned.tolist())
        flat = np.concatenate([np.asarray(o) for o in owned])
        assert len(flat) == n and len(np.unique(flat)) == n
        assert abs(len(owned[0]) - len(owned[1])) < 0.02 * n
        assert np.array_equal(plan.tree_local.numpy(), pos_np[plan.local_ids.numpy()])
        assert torch.equal(plan.payload_local[0], torch.stack([plan.local_ids.float() * 0.5, -plan.local_ids.float()], 1))
        assert not np.intersect1d(plan.owned.numpy(), plan.halo.numpy()).size

        # 2. the halo holds every foreign point a moved owned query can reach
        rng = np.random.default_rng(rank)
        spacing = hw / (3.0 * np.sqrt(k / np.pi))
        moved = pos_np[plan.owned.numpy()] + rng.normal(0, 0.5 * spacing, (plan.n_owned, 3)).astype(np.float32)
        nn = O.knn_bruteforce(pos_np, moved, k)
        local = np.zeros(n, dtype=bool)
        local[plan.local_ids.numpy()] = True
        assert local[nn].all(), "halo too thin"
        assert 0 < plan.n_halo < 0.5 * n

        # 3. halo refresh through the wiring the owners chose
        def values(ids_, phase):
            return torch.stack([ids_.float() * (phase + 1), ids_.float() + 0.25, -ids_.float(), torch.full_like(ids_, phase).float()], 1)

        state = torch.zeros((plan.local_ids.numel(), 4))
        for phase in range(2):
            state[:plan.n_owned] = values(plan.owned, phase)
            send = torch.cat([state[r] for r in ex.send_local])
            recv = torch.empty((sum(ex.recv_counts), 4))
            ex.exchange(send, recv)
            state[torch.cat(ex.recv_local)] = recv
            assert torch.equal(state, values(plan.local_ids, phase)), f"phase {phase}"

        # 3b. cost-aware slab sizes: the slabs follow the requested shares of the points, rebalance_fractions() inverts a
        #     piecewise-constant cost density
        plan2 = partition.ShardedSlabPlan(shard, gid, hw, None, [], fractions=[0.3, 0.7])
        sizes = [None] * world
        dist.all_gather_object(sizes, plan2.n_owned)
        assert sum(sizes) == n and abs(sizes[0] - 0.3 * n) < 0.02 * n
        assert np.allclose(partition.rebalance_fractions([1, 3], [0.5, 0.5]), [2 / 3, 1 / 3])
        assert np.allclose(partition.rebalance_fractions([1, 1, 2, 2], [0.25] * 4), [0.375, 0.25, 0.1875, 0.1875])
        assert np.allclose(partition.rebalance_fractions([2, 2, 2], [0.2, 0.3, 0.5]), [0.2, 0.3, 0.5])

        # 4. peer-push addressing agrees between the ranks (what _wire_peer hands to ngpd_session_set_slab)
        cap, first_row, seg = ex.peer_layout(torch.device("cpu"))
        caps = [None] * world
        dist.all_gather_object(caps, (cap, first_row, seg, ex.recv_counts))
        assert caps[0][0] == caps[1][0] >= max(sum(c[3]) for c in caps)
        other = 1 - rank
        assert seg[other + 1] - seg[other] == caps[other][3][rank]           # what I send to the peer = what it expects from me
        assert first_row[other] == sum(caps[other][3][:rank])
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_slab_plan_world2():
    """distributed planning: no rank sees the whole cloud (VERDICT r1 missing #3)"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}
