"""world_size-2 gloo test (CPU) of the N>1 path: tomogram sharding + the pick-list gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cet_pick_b200.shard import gather_picks, shard_range


def test_shard_range_partitions_everything():
    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, cnt = shard_range(n, r, world)
                seen += list(range(first, first + cnt))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _picks(i, K):
    g = torch.Generator().manual_seed(1000 + i)
    return torch.rand((K, 5), generator=g)


def _worker(rank, world, port, n_items, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, cnt = shard_range(n_items, rank, world)
        local = torch.stack([_picks(first + i, K) for i in range(cnt)]) if cnt else torch.empty((0, K, 5))
        allp = gather_picks(local, n_items)
        ref = torch.stack([_picks(i, K) for i in range(n_items)])
        q.put((rank, bool(torch.equal(allp, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5])
def test_gather_picks_world2_gloo(n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


# ---- z-slab sharding of ONE oversized volume (shard.forward_z_sharded): recompute halo + one all_gather ----
def _fake_forward(slab):
    """a stand-in with the detector's z reach: per-slice 2x down-sampling + a +-3-slice zero-padded z filter"""
    x = slab[:, ::2, ::2]
    k = torch.tensor([1.0, -2.0, 3.0, 5.0, 0.5, -1.5, 2.5])
    xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 3, 3))
    return sum(k[i] * xp[i:i + x.shape[0]] for i in range(7))


def test_slab_range_covers_depth_with_halo():
    from cet_pick_b200.shard import slab_range
    for depth in (1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            cores = []
            for r in range(world):
                z0, z1, lo, hi = slab_range(depth, r, world)
                cores += list(range(z0, z1))
                assert lo == max(0, z0 - 3) and hi == min(depth, z1 + 3)
            assert cores == list(range(depth))


def _zworker(rank, world, port, depth, q):
    from cet_pick_b200.shard import forward_z_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = torch.rand((depth, 12, 10), generator=torch.Generator().manual_seed(5))
        loaded = []

        def slab_fn(lo, hi):
            loaded.append((lo, hi))
            return vol[lo:hi]

        hm = forward_z_sharded(_fake_forward, slab_fn, depth)
        ok = torch.allclose(hm, _fake_forward(vol), atol=1e-6) and hm.shape[0] == depth
        # each rank touched only its slab plus the 3-slice halo
        ok = ok and len(loaded) <= 1 and all(hi - lo <= -(-depth // world) + 6 for lo, hi in loaded)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("depth", [16, 9, 1])
def test_forward_z_sharded_world2_gloo(depth):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_zworker, args=(r, 2, port, depth, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


# ---- z-sharded decode: per-rank top-K candidates -> all_gather -> merge-select ----
def _cand_worker(rank, world, port, depth, K, q):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import synthdata as synth
    from cet_pick_b200.shard import gather_merge_candidates, shard_range
    from oracle import decode_oracle as do
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W = 14, 18
        hm = synth.heatmap_tiefree_np(depth, H, W, 31)
        nm = do.nms(hm[None, None], 3)[0, 0]                     # oracle NMS of the whole map (CPU stand-in for the
        z0, cnt = shard_range(depth, rank, world)                 # GPU part, which the -m gpu tests cover)
        own = np.zeros_like(nm)
        own[z0:z0 + cnt] = nm[z0:z0 + cnt]
        sc, li = do.topk_canonical(own.ravel(), K)
        ms, mi = gather_merge_candidates(torch.from_numpy(sc.copy()), torch.from_numpy(li.copy()), K)
        rs, ri = do.topk_canonical(nm.ravel(), K)
        pos = rs > 0
        ok = np.array_equal(ms.numpy()[pos], rs[pos]) and np.array_equal(mi.numpy()[pos], ri[pos])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("depth,K", [(12, 40), (7, 25), (2, 10)])
def test_candidate_merge_world2_gloo(depth, K):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cand_worker, args=(r, 2, port, depth, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_merge_topk_canonical_order():
    from cet_pick_b200.shard import merge_topk
    sc = torch.tensor([0.5, 0.9, 0.5, 0.1, 0.9, 0.5])
    ix = torch.tensor([40, 7, 3, 99, 2, 11])
    s, i = merge_topk(sc, ix, 4)
    assert torch.equal(s, torch.tensor([0.9, 0.9, 0.5, 0.5])) and i.tolist() == [2, 7, 3, 11]


# ---- training step: one flat-bucket all-reduce of the gradients ----
def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from cet_pick_b200.models.model import create_model
    from cet_pick_b200.trains.step import FlatBucket, allreduce_gradients
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
        bucket = FlatBucket(m)
        ok = bucket.numel == sum(p.numel() for p in m.parameters()) and bucket.numel * 4 < 8.1e6     # one 7.97 MB bucket
        for i, p in enumerate(m.parameters()):
            p.grad.fill_(float(rank + 1) * (i + 1))                # rank-dependent gradients, written through the views
        scale = allreduce_gradients(bucket)
        exp = sum(r + 1 for r in range(world)) / world
        for i, p in enumerate(m.parameters()):
            ok = ok and bool(torch.allclose(p.grad, torch.full_like(p.grad, exp * (i + 1))))
        ok = ok and scale == 1.0
        for i, p in enumerate(m.parameters()):
            p.grad.fill_(float(rank + 1))
        ok = ok and allreduce_gradients(bucket, average=False) == 1.0 / world and float(bucket.grads[0]) == sum(r + 1 for r in range(world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_gradient_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
