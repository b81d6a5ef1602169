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
