"""Bank-sharded memory read over N GPUs (BASELINE configs[4]): parity against the single-GPU read + timing.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--n 50000 --k 32]
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tracking-anything-with-deva_b200'))
from deva import _native as nat  # noqa: E402
from deva.inference.sharded_memory import ShardedBankReader, shard_bounds  # noqa: E402

CK, CV = 64, 512


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--slots', dest='n', type=int, default=50000)
    ap.add_argument('--queries', dest='q', type=int, default=8160)
    ap.add_argument('--objects', dest='k', type=int, default=32)
    ap.add_argument('--check', action='store_true', help='compare with the unsharded read on rank 0 (needs the memory)')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--scatter', action='store_true', help='also run the fused readout + reduce-scatter-by-object '
                    '(peer-memory red.add from the GEMM epilogue) and compare it with the all-reduce read')
    a = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nat.require_device()
    g = torch.Generator(device=dev).manual_seed(0)  # same stream on every rank -> replicated inputs
    mk = torch.randn(CK, a.n, device=dev, generator=g)
    ms = 1 + torch.rand(a.n, device=dev, generator=g)
    qk = torch.randn(CK, a.q, device=dev, generator=g)
    qe = torch.sigmoid(torch.randn(CK, a.q, device=dev, generator=g))
    lo, hi = shard_bounds(a.n, world, rank)
    gv = torch.Generator(device=dev).manual_seed(1)
    # values are generated per shard from a seed that depends on the slot range only
    mv_full = None
    if a.check:
        mv_full = (torch.randn(a.k * CV, a.n, device=dev, generator=gv) if a.k * CV * a.n * 4 < 30e9 else None)
    mv = mv_full[:, lo:hi].contiguous() if mv_full is not None else torch.randn(a.k * CV, hi - lo, device=dev, generator=gv)
    rd = ShardedBankReader(CK, CV, a.k, hi - lo, lo, dev)
    rd.load(mk[:, lo:hi], ms[lo:hi], mv)
    del mv
    out = rd.read(qk, qe)
    torch.cuda.synchronize()
    res = {'world': world, 'n': a.n, 'q': a.q, 'k': a.k, 'shard': [lo, hi]}
    if a.check and mv_full is not None:
        ref = ShardedBankReader(CK, CV, a.k, a.n, 0, dev, group=None)
        ref.world, ref.rank = 1, 0
        ref.load(mk, ms, mv_full)
        want = ref.read(qk, qe)
        torch.cuda.synchronize()
        res['max_abs_diff_vs_unsharded'] = float((out - want).abs().max())
        res['scale'] = float(want.abs().max())
        del ref, want
    if a.scatter:
        own = rd.read_scatter(qk, qe, count_usage=False).clone()
        olo, ohi = rd.objects_of(rank)
        diff = (own - out[olo * CV:ohi * CV]).abs().max() if ohi > olo else torch.zeros((), device=dev)
        if world > 1:
            dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        res['scatter_max_abs_diff_vs_allreduce'] = float(diff)
        tss = []
        for _ in range(a.iters):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rd.read_scatter(qk, qe, count_usage=False); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tss.append(float(t))
        tss.sort()
        res['ms_scatter'] = tss[len(tss) // 2]
    ts = []
    for _ in range(a.iters):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rd.read(qk, qe); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    ts.sort()
    ms_read = ts[len(ts) // 2]
    flops = 2.0 * a.n * a.q * 2 * CK + 2.0 * a.k * CV * a.n * a.q
    res.update(ms=ms_read, dense_equiv_tflops=flops / ms_read / 1e9)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
