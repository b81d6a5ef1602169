"""GPU parity: C-ABI memory-read kernels vs the CPU oracle (oracle/memory_math.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import memory_math as mm

pytestmark = pytest.mark.gpu
CK = 64


def _native():
    from deva import _native
    _native.require_device()
    return _native


def _make(n, q, k_obj, cv, seed, lead=0):
    g = torch.Generator().manual_seed(seed)
    mk = torch.randn(CK, n, generator=g)
    ms = 1 + torch.rand(n, generator=g)
    qk = torch.randn(CK, q, generator=g)
    qe = torch.sigmoid(torch.randn(CK, q, generator=g))
    mv = torch.randn(k_obj * cv, n, generator=g)
    return mk, ms, qk, qe, mv


class Bank:
    """Minimal packed bank for the kernel tests (the product's bank lives in deva.inference)."""
    def __init__(self, nat, mk, ms, mv, lead=0):
        dev = 'cuda'
        n = mk.shape[1]
        self.n, self.lead = n, lead
        nw = n + lead
        self.nw = nw
        self.k_hi = torch.zeros(nw, 2 * CK, dtype=torch.float16, device=dev)
        self.k_lo = torch.zeros_like(self.k_hi)
        self.neg_s = torch.zeros(nw, device=dev)
        self.raw_key = torch.zeros(nw, CK, device=dev)
        self.raw_shr = torch.zeros(nw, device=dev)
        mk_d, ms_d = mk.to(dev).contiguous(), ms.to(dev).contiguous()
        nat.pack_keys(mk_d, None, n, 1, ms_d, CK, n, self.k_hi[lead:], self.k_lo[lead:], self.neg_s[lead:],
                      self.raw_key[lead:], None, self.raw_shr[lead:])
        rows = mv.shape[0]
        self.ld = (nw + 7) // 8 * 8 + 64
        self.values = torch.zeros(rows, self.ld, dtype=torch.float16, device=dev)
        nat.append_values(mv.to(dev).contiguous(), n, self.values[:, lead:], self.ld, rows, n)
        self.use = torch.zeros(nw, device=dev)
        self.life = torch.zeros(nw, device=dev) + 1e-7


def _read(nat, bank, qk, qe, k_obj, cv, top_k=30):
    dev = 'cuda'
    q = qk.shape[1]
    q_hi = torch.empty(q, 2 * CK, dtype=torch.float16, device=dev)
    q_lo = torch.empty_like(q_hi)
    bsq = torch.empty(q, device=dev)
    qk_d, qe_d = qk.to(dev).contiguous(), qe.to(dev).contiguous()
    nat.pack_query(qk_d, qe_d, q, 1, CK, q, q_hi, q_lo, bsq)
    ws = torch.empty(nat.simtopk_workspace_bytes(q), dtype=torch.uint8, device=dev)
    idx = torch.empty(q, 32, dtype=torch.int32, device=dev)
    w = torch.empty(q, 32, device=dev)
    ldp = (bank.nw + 7) // 8 * 8
    P = torch.empty(q, ldp, dtype=torch.float16, device=dev)
    nat.sim_topk(bank.k_hi, bank.k_lo, bank.neg_s, bank.nw, bank.lead, q_hi, q_lo, bsq, q, CK, top_k, ws, idx, w, P,
                 ldp, bank.use, bank.life, bank.lead, False, True)
    out = torch.empty(k_obj * cv, q, device=dev)
    nat.readout(bank.values, bank.ld, k_obj * cv, [i * cv for i in range(k_obj)], [i * cv for i in range(k_obj)], cv,
                P, ldp, bank.nw, q, out, q)
    # fused variant: affinity tiles generated on chip from (idx, w) - must equal the dense-operand GEMM
    out_sp = torch.empty_like(out)
    rws = torch.empty(nat.readout_sparse_workspace_bytes(q, bank.nw), dtype=torch.uint8, device=dev)
    nat.readout_sparse(bank.values, bank.ld, k_obj * cv, [i * cv for i in range(k_obj)], [i * cv for i in range(k_obj)],
                       cv, idx, w, top_k, bank.nw, q, rws, out_sp, q)
    tok = torch.empty(k_obj, q, cv, dtype=torch.float16, device=dev)
    nat.readout_sparse(bank.values, bank.ld, k_obj * cv, [i * cv for i in range(k_obj)], [i * cv for i in range(k_obj)],
                       cv, idx, w, top_k, bank.nw, q, rws, None, 0, out_tok=tok)
    torch.cuda.synchronize()
    scale = max(1.0, float(out.abs().max()))
    assert float((out_sp - out).abs().max()) < 1e-5 * scale, 'sparse-affinity readout differs from the dense-operand GEMM'
    assert float((tok.float().permute(0, 2, 1).reshape(k_obj * cv, q) - out).abs().max()) < 2e-3 * scale
    return idx.cpu(), w.cpu(), P.float().cpu(), out.cpu(), (q_hi, q_lo, bsq)


def _check_read(n, q, k_obj, cv, seed, lead=0, top_k=30):
    nat = _native()
    mk, ms, qk, qe, mv = _make(n, q, k_obj, cv, seed)
    bank = Bank(nat, mk, ms, mv, lead)
    idx, w, P, out, _ = _read(nat, bank, qk, qe, k_obj, cv, top_k)
    sim = mm.similarity(mk.double(), ms.double(), qk.double(), qe.double())
    ref_idx, ref_w = mm.topk_softmax(sim, top_k)  # [k, Q]
    idx = idx[:, :top_k].long() - lead
    assert int(idx.min()) >= 0 and int(idx.max()) < n
    # membership: identical sets except where the k-th / (k+1)-th similarities are within 2e-4
    srt = torch.sort(sim, 0, descending=True)[0]
    gap = (srt[top_k - 1] - srt[top_k]).abs() if n > top_k else torch.full((q, ), 1.0, dtype=sim.dtype)
    same = torch.tensor([set(idx[i].tolist()) == set(ref_idx[:, i].tolist()) for i in range(q)])
    assert bool((same | (gap < 2e-4)).all()), f'top-k sets differ on {int((~same).sum())} queries'
    # values at the selected slots match fp64 similarity to ~fp32 accuracy
    sel_sim = torch.gather(sim.t(), 1, idx)  # [Q, k]
    e = torch.exp(sel_sim - sel_sim.max(1, keepdim=True)[0])
    w_ref = (e / e.sum(1, keepdim=True)).float()
    assert float((w[:, :top_k] - w_ref).abs().max()) < 2e-5
    assert bool((w[:, top_k:] == 0).all())
    # sorted by descending similarity
    assert bool((sel_sim[:, :-1] - sel_sim[:, 1:] > -2e-4).all())
    # dense affinity rows and usage
    aff = torch.zeros(q, bank.nw).scatter_(1, idx + lead, w[:, :top_k])
    assert float((P[:, :bank.nw] - aff).abs().max()) < 6e-4  # fp16 rounding of weights in [0,1]
    usage = bank.use.cpu()
    assert float((usage - aff.sum(0)).abs().max()) < 1e-4
    assert float((bank.life.cpu()[lead:] - 1.0).abs().max()) < 1e-5
    # readout: fp16 operands, fp32 accumulate
    big = n * q * mv.shape[0] > 2e11  # C3-size problem: evaluate the oracle's readout through its k non-zeros per query
    if big:
        ref_out = _sparse_readout(mv.double(), ref_idx.t(), ref_w.t()).float()
    else:
        ref_out = mm.readout(mm.dense_affinity(sim, top_k), mv.double()).float()
    err = float((out - ref_out).abs().max())
    scale = float(ref_out.abs().max())
    assert err < 4e-3 * max(1.0, scale), (err, scale)
    if bool(same.all()):
        # against the same fp16-rounded operands the GEMM must be fp32-exact
        if big:
            ref16 = _sparse_readout(mv.half().double(), idx, w[:, :top_k].half().double()).float()
        else:
            ref16 = (mv.half().double() @ P[:, lead:lead + n].double().t()).float()
        assert float((out - ref16).abs().max()) < 2e-4 * max(1.0, scale)
    return err


def _sparse_readout(mv, idx, w, chunk=256):
    """oracle.memory_math.readout for an affinity with k non-zeros per query: out[:, q] = sum_j w[q, j] * mv[:, idx[q, j]]
    (the same sum as the dense product affinity^T . values, memory_manager.py:64-75, without the zero terms)."""
    q = idx.shape[0]
    out = torch.empty(mv.shape[0], q, dtype=mv.dtype)
    for a in range(0, q, chunk):
        cols = mv[:, idx[a:a + chunk].reshape(-1)].view(mv.shape[0], -1, idx.shape[1])  # [R, chunk, k]
        out[:, a:a + chunk] = (cols * w[a:a + chunk].to(mv.dtype).unsqueeze(0)).sum(-1)
    return out


def test_operand_split_is_fp32_accurate():
    nat = _native()
    mk, ms, qk, qe, mv = _make(300, 50, 1, 128, 3)
    bank = Bank(nat, mk, ms, mv)
    s = ms / 8.0
    want = torch.cat([(s * mk * mk).t(), (s * mk).t()], 1)
    got = bank.k_hi.float().cpu().double() + bank.k_lo.float().cpu().double()
    assert float((got - want.double()).abs().max()) < 2e-6 * float(want.abs().max())
    assert torch.equal(bank.raw_key.cpu(), mk.t().contiguous())
    assert float((bank.neg_s.cpu() + s).abs().max()) < 1e-6


def test_golden_fixture(golden_dir):
    """The reference-minted fixture, through the CUDA path."""
    nat = _native()
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, 'memory_read.npz')).items()}
    mk, ms, qk, qe = g['mk'], g['ms'].reshape(-1), g['qk'], g['qe']
    mv = torch.zeros(256, mk.shape[1])
    mv[:32] = g['mv'][0]
    mv[128:160] = g['mv'][1]
    bank = Bank(nat, mk, ms, mv)
    idx, w, P, out, _ = _read(nat, bank, qk, qe, 2, 128)
    ref_idx = g['topk_idx'].t()
    assert all(set(idx[i, :30].tolist()) == set(ref_idx[i].tolist()) for i in range(idx.shape[0]))
    assert float((bank.use.cpu() - g['usage']).abs().max()) < 1e-4
    got = torch.cat([out[:32], out[128:160]]).view(2, 32, -1)
    assert float((got - g['readout']).abs().max()) < 4e-3


@pytest.mark.parametrize('n,q,k_obj,cv,lead', [
    (30, 5, 1, 128, 0),          # N == top_k, one partial tile
    (257, 37, 1, 128, 3),        # ragged everything, masked lead slots
    (1000, 300, 2, 256, 5),
    (2000, 1620, 5, 512, 0),     # BASELINE config 2
    (10000, 8160, 16, 512, 0),   # BASELINE config 3: the headline shape (1080p, 16 objects, 10k slots)
])
def test_read_matches_oracle(n, q, k_obj, cv, lead):
    _check_read(n, q, k_obj, cv, seed=n + q, lead=lead)


def test_small_topk():
    _check_read(500, 64, 1, 128, seed=9, top_k=7)


def test_dense_softmax_matches_oracle():
    nat = _native()
    n, q = 900, 128
    mk, ms, qk, qe, mv = _make(n, q, 1, 128, 21)
    bank = Bank(nat, mk, ms, mv, lead=2)
    dev = 'cuda'
    q_hi = torch.empty(q, 2 * CK, dtype=torch.float16, device=dev)
    q_lo = torch.empty_like(q_hi)
    bsq = torch.empty(q, device=dev)
    nat.pack_query(qk.to(dev).contiguous(), qe.to(dev).contiguous(), q, 1, CK, q, q_hi, q_lo, bsq)
    ld = (bank.nw + 7) // 8 * 8
    sim_ws = torch.empty(q, ld, device=dev)
    P = torch.zeros(q, ld, dtype=torch.float16, device=dev)
    shr_out = torch.empty(q, device=dev)
    nat.sim_dense_softmax(bank.k_hi, bank.k_lo, bank.neg_s, bank.raw_shr, bank.nw, bank.lead, q_hi, q_lo, bsq, q, CK,
                          sim_ws, ld, P, ld, shr_out)
    torch.cuda.synchronize()
    sim = mm.similarity(mk, ms, qk, qe)
    assert float((sim_ws.cpu()[:, 2:2 + n] - sim.t()).abs().max()) < 1e-4
    aff = mm.dense_affinity(sim, None)
    assert float((P.float().cpu()[:, 2:2 + n] - aff.t()).abs().max()) < 6e-4
    assert float((shr_out.cpu() - (ms.reshape(1, -1) @ aff).reshape(-1)).abs().max()) < 1e-4


def test_sharded_reader_single_rank_matches_oracle():
    """ShardedBankReader (world 1): local top-k -> list merge -> localise -> sparse readout with zero-weight padding."""
    nat = _native()
    from deva.inference.sharded_memory import ShardedBankReader, localise, shard_bounds
    n, q, k_obj, cv = 3000, 500, 2, 128
    mk, ms, qk, qe, mv = _make(n, q, k_obj, cv, 77)
    rd = ShardedBankReader(CK, cv, k_obj, n, 0, 'cuda')
    rd.load(mk.cuda(), ms.cuda(), mv.cuda())
    out = rd.read(qk.cuda(), qe.cuda()).cpu()
    sim = mm.similarity(mk.double(), ms.double(), qk.double(), qe.double())
    ref = mm.readout(mm.dense_affinity(sim, 30), mv.double()).float()
    assert float((out - ref).abs().max()) < 4e-3 * max(1.0, float(ref.abs().max()))
    assert float((rd.use_cnt.cpu() - mm.dense_affinity(sim, 30).sum(1).float()).abs().max()) < 1e-3
    # emulate 3 shards on one device: per-shard partial readouts must add up to the full one
    total = torch.zeros_like(out)
    lists_v, lists_i = [], []
    shards = [shard_bounds(n, 3, r) for r in range(3)]
    readers = []
    for lo, hi in shards:
        r = ShardedBankReader(CK, cv, k_obj, hi - lo, lo, 'cuda')
        r.load(mk[:, lo:hi].cuda(), ms[lo:hi].cuda(), mv[:, lo:hi].cuda())
        readers.append(r)
    # run the protocol by hand (no process group): gather the local lists, merge, localise, read
    qk_d, qe_d = qk.cuda(), qe.cuda()
    q_hi = torch.empty(q, 2 * CK, dtype=torch.float16, device='cuda'); q_lo = torch.empty_like(q_hi)
    bsq = torch.empty(q, device='cuda')
    nat.pack_query(qk_d.contiguous(), qe_d.contiguous(), q, 1, CK, q, q_hi, q_lo, bsq)
    for r in readers:
        idx = torch.empty(q, 32, dtype=torch.int32, device='cuda'); w = torch.empty(q, 32, device='cuda')
        s_ = torch.empty(q, 32, device='cuda')
        ws = torch.empty(nat.simtopk_workspace_bytes(q), dtype=torch.uint8, device='cuda')
        nat.sim_topk(r.k_hi, r.k_lo, r.neg_s, r.n, 0, q_hi, q_lo, bsq, q, CK, 30, ws, idx, w, None, 0, None, None, 0,
                     False, False, out_sim=s_)
        valid = torch.arange(32, device='cuda').view(1, -1) < 30
        lists_v.append(s_.t().contiguous())
        lists_i.append(torch.where(valid, idx + r.offset, torch.full_like(idx, -1)).t().contiguous())
    g_sel = torch.empty(q, 32, dtype=torch.int32, device='cuda'); g_w = torch.empty(q, 32, device='cuda')
    nat.merge_lists(torch.stack(lists_v), torch.stack(lists_i), 3, 30, q, q, g_sel, g_w)
    for r in readers:
        il, wl = localise(g_sel, g_w, r.offset, r.offset + r.n)
        part = torch.zeros(k_obj * cv, q, device='cuda')
        rws = torch.empty(nat.readout_sparse_workspace_bytes(q, r.n), dtype=torch.uint8, device='cuda')
        rows = [i * cv for i in range(k_obj)]
        nat.readout_sparse(r.values, r.ld, k_obj * cv, rows, rows, cv, il, wl, 32, r.n, q, rws, part, q)
        total += part.cpu()
    torch.cuda.synchronize()
    assert float((total - out).abs().max()) < 1e-5 * max(1.0, float(out.abs().max()))


def test_scatter_readout_single_rank():
    """Fused readout + reduce-scatter by object, degenerate case of one rank: the red.add epilogue into the (local)
    owner buffer reproduces the stored readout; two reads in a row do not accumulate."""
    from deva import _native
    from deva.inference.sharded_memory import ShardedBankReader
    _native.require_device()
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(5)
    ck, cv, k, n, q = 64, 512, 3, 3000, 1000
    mk = torch.randn(ck, n, device=dev, generator=g)
    ms = 1 + torch.rand(n, device=dev, generator=g)
    mv = torch.randn(k * cv, n, device=dev, generator=g)
    qk = torch.randn(ck, q, device=dev, generator=g)
    qe = torch.sigmoid(torch.randn(ck, q, device=dev, generator=g))
    rd = ShardedBankReader(ck, cv, k, n, 0, dev)
    rd.load(mk, ms, mv)
    want = rd.read(qk, qe, count_usage=False)
    for _ in range(2):
        got = rd.read_scatter(qk, qe, count_usage=False)
        torch.cuda.synchronize()
        assert got.shape == want.shape
        assert float((got - want).abs().max()) < 1e-5 * float(want.abs().max()) + 1e-6


def test_training_time_read_memory_matches_reference(golden_dir):
    """DEVA.read_memory (reference network.py:72-92: get_affinity without top-k -> readout) through the
    similarity / row-softmax / dense readout kernels, against the unmodified reference's output on the same tensors."""
    from deva.model.network import DEVA
    g = {k: torch.from_numpy(v).cuda() for k, v in np.load(os.path.join(golden_dir, 'read_memory.npz')).items()}
    cfg = dict(key_dim=64, value_dim=g['mv'].shape[2], pix_feat_dim=512)
    net = DEVA(cfg).cuda().eval()
    out = net.read_memory(g['qk'], g['qe'], g['mk'], g['ms'], g['mv'])
    torch.cuda.synchronize()
    assert tuple(out.shape) == tuple(g['out'].shape)
    err = float((out - g['out']).abs().max())
    assert err < 2e-3 * max(1.0, float(g['out'].abs().max())), err  # fp16 operands of the readout GEMM, fp32 accumulate


def test_warm_started_topk_is_exact():
    """Temporal warm start (prev_idx): the slots selected for the previous frame's queries bound this frame's k-th best
    similarity from below; the streaming top-k then skips everything under the bound.  The result must be IDENTICAL to
    the cold read - also when the previous selection is stale, partly invalid or garbage."""
    nat = _native()
    n, q, cv = 6000, 700, 128
    mk, ms, qk, qe, mv = _make(n, q, 1, cv, 5)
    bank = Bank(nat, mk, ms, mv, lead=3)
    dev = 'cuda'

    def read(qk_, prev=None):
        q_hi = torch.empty(q, 2 * CK, dtype=torch.float16, device=dev)
        q_lo = torch.empty_like(q_hi)
        bsq = torch.empty(q, device=dev)
        nat.pack_query(qk_.to(dev).contiguous(), qe.to(dev).contiguous(), q, 1, CK, q, q_hi, q_lo, bsq)
        ws = torch.empty(nat.simtopk_workspace_bytes(q), dtype=torch.uint8, device=dev)
        idx = torch.empty(q, 32, dtype=torch.int32, device=dev) if prev is None else prev  # aliasing prev/out is allowed
        w = torch.empty(q, 32, device=dev)
        thr = torch.empty(q, device=dev)
        nat.sim_topk(bank.k_hi, bank.k_lo, bank.neg_s, bank.nw, bank.lead, q_hi, q_lo, bsq, q, CK, 30, ws, idx, w, None, 0,
                     None, None, 0, False, False, prev_idx=prev, thr_ws=thr)
        torch.cuda.synchronize()
        return idx.clone(), w.clone(), thr.clone()

    idx0, _, _ = read(qk)
    g = torch.Generator().manual_seed(9)
    qk1 = qk + 0.05 * torch.randn(qk.shape, generator=g)  # "next frame": most of the top-k persists
    cold_idx, cold_w, _ = read(qk1)
    warm_idx, warm_w, thr = read(qk1, prev=idx0.clone())
    assert torch.equal(warm_idx, cold_idx) and torch.equal(warm_w, cold_w)
    assert bool(torch.isfinite(thr).all())  # every query got a bound
    # the bound is tight: it sits close under the true k-th best similarity
    sim = mm.similarity(mk.double(), ms.double(), qk1.double(), qe.double())
    kth = torch.sort(sim, 0, descending=True)[0][29]
    assert bool((thr.cpu().double() <= kth + 1e-6).all()) and float((kth - thr.cpu().double()).median()) < 0.5
    # a very different query (stale selection), out-of-range and lead-slot indices, and pure garbage: still exact
    qk2 = torch.randn(qk.shape, generator=g)
    cold2_idx, cold2_w, _ = read(qk2)
    stale_idx, stale_w, _ = read(qk2, prev=idx0.clone())
    assert torch.equal(stale_idx, cold2_idx) and torch.equal(stale_w, cold2_w)
    junk = torch.randint(-5, bank.nw + 50, (q, 32), dtype=torch.int32, generator=g).to(dev)
    junk_idx, junk_w, _ = read(qk2, prev=junk)
    assert torch.equal(junk_idx, cold2_idx) and torch.equal(junk_w, cold2_w)
