"""world_size-2 gloo test of the clip-parallel host logic (runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from deva.utils.dist_utils import assign_clips, max_over_ranks, sum_over_ranks
    lengths = [40, 10, 10, 10, 30, 5, 5, 50]
    mine = assign_clips(len(lengths), world, rank, lengths)
    frames = sum(lengths[i] for i in mine)
    dist.barrier()
    slowest = max_over_ranks(1.0 + rank)
    total = sum_over_ranks(frames)
    # every clip is owned exactly once
    owned = [torch.zeros(len(lengths), dtype=torch.int64) for _ in range(world)]
    me = torch.zeros(len(lengths), dtype=torch.int64)
    me[mine] = 1
    dist.all_gather(owned, me)
    out[rank] = (mine, frames, slowest, total, torch.stack(owned).sum(0).tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_clip_sharding_two_ranks():
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, 'tracking-anything-with-deva_b200'))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (m0, f0, s0, t0, o0), (m1, f1, s1, t1, o1) = out[0], out[1]
    assert sorted(m0 + m1) == list(range(8)) and not set(m0) & set(m1)
    assert o0 == [1] * 8 and o1 == [1] * 8
    assert s0 == s1 == 2.0 and t0 == t1 == 160.0
    assert abs(f0 - f1) <= 10  # balanced by frame count


def test_round_robin_assignment():
    from deva.utils.dist_utils import assign_clips
    assert assign_clips(64, 8, 3) == list(range(3, 64, 8))
    assert sum(len(assign_clips(10, 4, r)) for r in range(4)) == 10


def test_shard_bounds_and_localise():
    from deva.inference.sharded_memory import localise, shard_bounds
    n, world = 50000, 8
    cover = [shard_bounds(n, world, r) for r in range(world)]
    assert cover[0][0] == 0 and cover[-1][1] == n
    assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    assert all(lo % 8 == 0 for lo, _ in cover)
    assert shard_bounds(10, 4, 3) == (10, 10)  # empty trailing shard
    idx = torch.tensor([[3, 17, 40, 9]], dtype=torch.int32)
    w = torch.tensor([[0.4, 0.3, 0.2, 0.1]])
    li, lw = localise(idx, w, 8, 24)
    assert li.tolist() == [[0, 9, 0, 1]] and lw.tolist() == [[0.0, 0.30000001192092896, 0.0, 0.10000000149011612]]
    # weights of all shards add back up to the global list
    tot = sum(localise(idx, w, *shard_bounds(48, 3, r))[1] for r in range(3))
    assert torch.allclose(tot, w)


def test_sharded_core_partitions():
    """token / object partitions of the bank-sharded core (deva.inference.sharded_core): disjoint, exhaustive, balanced."""
    from deva.inference.sharded_core import object_bounds, token_bounds
    for n in (30, 1620, 8160, 7):
        for world in (1, 2, 4, 8):
            cover = [token_bounds(n, world, r) for r in range(world)]
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cover]
            assert max(sizes) - min(sizes) <= 1
    for k in (0, 1, 2, 3, 16, 32, 33):
        for world in (1, 2, 4, 8):
            cover = [object_bounds(k, world, r) for r in range(world)]
            owned = [i for a, b in cover for i in range(a, b)]
            assert owned == list(range(k))
            per = -(-k // world) if k else 0
            assert all(b - a <= per for a, b in cover)


def _sharded_read_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import memory_math as mm
    from oracle import sharded
    torch.manual_seed(0)  # every rank draws the same bank and queries, then keeps its slice
    ck, cv, k_obj, n, q, top_k = 16, 8, 5, 203, 37, 30
    mem_key, mem_shr = torch.randn(ck, n, dtype=torch.float64), 1 + torch.rand(n, dtype=torch.float64)
    values = torch.randn(k_obj * cv, n, dtype=torch.float64)
    qk, qe = torch.randn(ck, q, dtype=torch.float64), torch.sigmoid(torch.randn(ck, q, dtype=torch.float64))
    lo, hi = sharded.token_bounds(n, world, rank)
    ro, usage, idx, w = sharded.sharded_read(mem_key[:, lo:hi], mem_shr[lo:hi], values[:, lo:hi], lo, qk, qe, top_k, k_obj)
    ref_ro, ref_usage, ref_idx, ref_w = mm.read(mem_key, mem_shr, qk, qe, values, top_k)
    per = -(-k_obj // world)
    a, b = min(k_obj, rank * per), min(k_obj, (rank + 1) * per)
    err_ro = float((ro - ref_ro[a * cv:b * cv]).abs().max()) if b > a else 0.0
    err_use = float((usage - ref_usage[lo:hi]).abs().max())
    same_set = bool((idx.sort(0)[0] == ref_idx.sort(0)[0]).all())
    err_w = float((w - ref_w).abs().max())
    # consolidation softmax over column slices of a [P, n] similarity
    sim = torch.randn(6, n, dtype=torch.float64) * 3
    pw = sharded.sharded_row_softmax(sim[:, lo:hi])
    err_sm = float((pw - torch.softmax(sim, dim=1)[:, lo:hi]).abs().max())
    out[rank] = (err_ro, err_use, same_set, err_w, err_sm, tuple(ro.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_bank_sharded_read_protocol_four_ranks():
    """The collective protocol of the bank-sharded read (local top-k -> all-gather -> global top-k + softmax -> partial
    read-out -> reduce-scatter by object) and of the sharded consolidation softmax, restated in oracle/sharded.py, equals the
    unsharded reference math at world size 4 with ragged token / object partitions (203 slots, 5 objects)."""
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    world = 4
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_read_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    shapes = []
    for r in range(world):
        err_ro, err_use, same_set, err_w, err_sm, shape = out[r]
        assert err_ro < 1e-12 and err_use < 1e-12 and same_set and err_w < 1e-12 and err_sm < 1e-12, (r, out[r])
        shapes.append(shape[0])
    assert shapes == [16, 16, 8, 0]  # 5 objects in blocks of 2: ranks own 2, 2, 1, 0 objects (CV = 8)
