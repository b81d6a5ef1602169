"""Micro-benchmark of the memory-read kernels at BASELINE configs (CUDA events, L2 flushed)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tracking-anything-with-deva_b200'))
from deva import _native as nat  # noqa: E402

CK, CV = 64, 512


def run(n, q, k_obj, iters=10, top_k=30):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    mk = torch.randn(CK, n, device=dev, generator=g)
    ms = 1 + torch.rand(n, device=dev, generator=g)
    qk = torch.randn(CK, q, device=dev, generator=g)
    qe = torch.sigmoid(torch.randn(CK, q, device=dev, generator=g))
    rows = k_obj * CV
    ld = (n + 7) // 8 * 8
    k_hi = torch.zeros(n, 2 * CK, dtype=torch.float16, device=dev); k_lo = torch.zeros_like(k_hi)
    neg_s = torch.zeros(n, device=dev); raw_key = torch.zeros(n, CK, device=dev); raw_shr = torch.zeros(n, device=dev)
    nat.pack_keys(mk, None, n, 1, ms, CK, n, k_hi, k_lo, neg_s, raw_key, None, raw_shr)
    values = torch.randn(rows, ld, device=dev, generator=g).half()
    q_hi = torch.empty(q, 2 * CK, dtype=torch.float16, device=dev); q_lo = torch.empty_like(q_hi)
    bsq = torch.empty(q, device=dev)
    ws = torch.empty(nat.simtopk_workspace_bytes(q), dtype=torch.uint8, device=dev)
    idx = torch.empty(q, 32, dtype=torch.int32, device=dev); w = torch.empty(q, 32, device=dev)
    P = torch.empty(q, ld, dtype=torch.float16, device=dev)
    use = torch.zeros(n, device=dev); life = torch.zeros(n, device=dev)
    out = torch.empty(rows, q, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    vr = [i * CV for i in range(k_obj)]

    def stage_pack():
        nat.pack_query(qk, qe, q, 1, CK, q, q_hi, q_lo, bsq)

    def stage_topk():
        nat.sim_topk(k_hi, k_lo, neg_s, n, 0, q_hi, q_lo, bsq, q, CK, top_k, ws, idx, w, P, ld, use, life, 0, False, True)

    def stage_read():
        nat.readout(values, ld, rows, vr, vr, CV, P, ld, n, q, out, q)

    rws = torch.empty(nat.readout_sparse_workspace_bytes(q, n), dtype=torch.uint8, device=dev)

    def stage_topk_nodense():
        nat.sim_topk(k_hi, k_lo, neg_s, n, 0, q_hi, q_lo, bsq, q, CK, top_k, ws, idx, w, None, 0, use, life, 0, False, True)

    def stage_read_sparse():
        nat.readout_sparse(values, ld, rows, vr, vr, CV, idx, w, top_k, n, q, rws, out, q)

    res = {}
    for name, fn in (('pack_query', stage_pack), ('sim_topk', stage_topk), ('readout', stage_read),
                     ('sim_topk_lists_only', stage_topk_nodense), ('readout_fused', stage_read_sparse)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[name] = ts[len(ts) // 2]
    flops = 2.0 * n * q * 2 * CK + 2.0 * rows * n * q
    tot = res['sim_topk_lists_only'] + res['readout_fused']
    res.update(n=n, q=q, k_obj=k_obj, gflop=flops / 1e9, fused_tflops=flops / tot / 1e9,
               readout_tflops=2.0 * rows * n * q / res['readout'] / 1e9,
               readout_fused_tflops=2.0 * rows * n * q / res['readout_fused'] / 1e9)
    return res


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c2,c3')
    a = ap.parse_args()
    nat.require_device()
    cfgs = {'c2': (2000, 1620, 5), 'c3': (10000, 8160, 16), 'c5': (6250, 8160, 32)}
    for c in a.configs.split(','):
        print(json.dumps({c: run(*cfgs[c])}))
