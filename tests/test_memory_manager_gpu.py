"""GPU parity of MemoryManager (bank bookkeeping + reads) against oracle.memory_bank.MemoryOracle on random streams."""
import pytest
import torch

from oracle.memory_bank import MemoryOracle

pytestmark = pytest.mark.gpu

CFG = dict(key_dim=64, value_dim=128, pix_feat_dim=512, mem_every=1, enable_long_term=True, chunk_size=-1, top_k=30,
           enable_long_term_count_usage=True, max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=128,
           max_long_term_elements=300)


def _frame(g, k, h, w):
    key = torch.randn(1, 64, h, w, generator=g)
    # nearly constant shrinkage: with a wide spread the ranking is slot-dominated, many slots never enter any top-k,
    # and the reference's topk(usage) prototype choice is then decided by arbitrary tie-breaking among zeros
    shr = 1 + 0.05 * torch.rand(1, 1, h, w, generator=g)
    sel = torch.sigmoid(torch.randn(1, 64, h, w, generator=g))
    val = torch.randn(1, k, 128, h, w, generator=g)
    return key, shr, sel, val


def _compare(mm, ref, key, sel, tol):
    got = mm.match_memory(key.cuda(), sel.cuda())
    want = ref.read(key, sel)
    assert set(got.keys()) == set(want.keys())
    worst = 0.0
    for o in want:
        worst = max(worst, float((got[o].float().cpu() - want[o]).abs().max()))
    assert worst < tol, worst
    return worst


@pytest.mark.parametrize('layout', ['nchw', 'nhwc'])
def test_stream_with_consolidation_eviction_new_bucket_and_purge(layout):
    from deva.inference.memory_manager import MemoryManager
    g = torch.Generator().manual_seed(4)
    h, w = 6, 8  # 48 tokens per frame: consolidation at 288, candidates 144 >= 128 prototypes
    mm, ref = MemoryManager(CFG), MemoryOracle(CFG)
    mm.readout_layout = layout
    objs = [3, 9]
    for t in range(26):
        if t == 7:
            objs = objs + [21]           # new object -> second bucket
        if t == 15:                      # object 9 disappears: purge (incorporate_detection path)
            objs = [3, 21]
            mm.purge_except(objs)
            ref.keep_only(objs)
        key, shr, sel, val = _frame(g, len(objs), h, w)
        if 0 < t <= 16:
            # Every kind of bank event has happened by t = 16 (append, consolidation, second bucket, eviction, purge).
            # Values are stored in fp16 and prototypes are re-read from them: allow fp16-level deviation.  Later
            # steps only check the bookkeeping: prototype choice (topk of usage) and eviction (strict threshold) are
            # discontinuous in the usage counters, so fp32-level differences eventually pick different slots.
            _compare(mm, ref, key, sel, 5e-3)
        elif t > 16:
            got = mm.match_memory(key.cuda(), sel.cuda())
            ref.read(key, sel)
            assert all(bool(torch.isfinite(v).all()) for v in got.values())
        mm.add_memory(key.cuda(), shr.cuda(), val.cuda(), list(objs), selection=sel.cuda())
        ref.add(key, shr, val, list(objs), selection=sel)
        sizes = {b: (mm.work_mem.size(b), mm.long_mem.size(b)) for b in mm.work_mem.buckets}
        assert sizes == ref.sizes(), (t, sizes, ref.sizes())
        assert mm.work_mem.buckets == {b: r.objects for b, r in ref.work.buckets.items()}
    for b in ref.work.buckets:
        got = mm.work_mem.get_usage(b).cpu()
        assert got.shape == ref.work.usage(b).shape and bool(torch.isfinite(got).all())
    # reference-shaped views
    b0 = next(iter(mm.work_mem.buckets))
    assert tuple(mm.work_mem.key[b0].shape) == (64, mm.work_mem.size(b0))
    assert tuple(mm.work_mem.shrinkage[b0].shape) == (1, mm.work_mem.size(b0))
    assert tuple(mm.work_mem.value[3].shape) == (128, mm.work_mem.size(b0))
    torch.testing.assert_close(mm.work_mem.key[b0].cpu(), ref.work.buckets[b0].key)  # working keys are exact copies
    assert mm.work_mem.num_objects == 2 and 3 in mm.work_mem and 9 not in mm.work_mem


def test_long_term_disabled_grows_and_reads():
    from deva.inference.memory_manager import MemoryManager
    cfg = dict(CFG, enable_long_term=False, enable_long_term_count_usage=False)
    g = torch.Generator().manual_seed(5)
    mm, ref = MemoryManager(cfg), MemoryOracle(cfg)
    for t in range(20):  # 20 * 35 tokens > the initial 16-frame capacity -> exercises the bank growth path
        key, shr, sel, val = _frame(g, 1, 5, 7)
        if t > 0:
            _compare(mm, ref, key, sel, 5e-3)
        mm.add_memory(key.cuda(), shr.cuda(), val.cuda(), [1], selection=sel.cuda())
        ref.add(key, shr, val, [1], selection=sel)
    assert mm.work_mem.size(0) == 20 * 35


def test_topk_larger_than_bank_raises_like_the_reference():
    from deva.inference.memory_manager import MemoryManager
    g = torch.Generator().manual_seed(6)
    mm = MemoryManager(CFG)
    key, shr, sel, val = _frame(g, 1, 4, 5)  # 20 tokens < top_k = 30 -> torch.topk raises in the reference
    mm.add_memory(key.cuda(), shr.cuda(), val.cuda(), [1], selection=sel.cuda())
    with pytest.raises(RuntimeError):
        mm.match_memory(key.cuda(), sel.cuda())


def test_memory_utils_api():
    from deva.model import memory_utils as mu
    from oracle import memory_math as om
    g = torch.Generator().manual_seed(7)
    mk, ms = torch.randn(1, 64, 5, 9, generator=g), 1 + torch.rand(1, 1, 5, 9, generator=g)
    qk, qe = torch.randn(1, 64, 4, 6, generator=g), torch.sigmoid(torch.randn(1, 64, 4, 6, generator=g))
    sim = mu.get_similarity(mk.cuda(), ms.cuda(), qk.cuda(), qe.cuda())
    want = om.similarity(mk[0].flatten(1), ms[0].flatten(), qk[0].flatten(1), qe[0].flatten(1))
    assert tuple(sim.shape) == (1, 45, 24)
    assert float((sim[0].cpu() - want).abs().max()) < 1e-4
    aff = mu.do_softmax(sim, top_k=30)
    assert float((aff[0].cpu() - om.dense_affinity(want, 30)).abs().max()) < 1e-4


def test_incorporate_detection_smoke(synthetic_sd):
    """Detection merge path: new objects enter, an unmatched object is purged after too many misses."""
    from deva.inference.inference_core import DEVAInferenceCore
    from deva.inference.object_info import ObjectInfo
    from deva.model.network import DEVA
    cfg = dict(CFG, value_dim=512, mem_every=5, max_mid_term_frames=10, min_mid_term_frames=5,
               max_long_term_elements=10000, max_missed_detection_count=1, max_num_objects=-1)
    net = DEVA(cfg).cuda().eval()
    net.load_weights({k: v.cuda() for k, v in synthetic_sd.items()})
    core = DEVAInferenceCore(net, cfg)
    g = torch.Generator().manual_seed(8)
    H, W = 96, 128
    img = torch.randn(3, H, W, generator=g).cuda()
    det = torch.zeros(H, W, dtype=torch.long)
    det[8:40, 8:60] = 5
    det[50:90, 60:120] = 6
    p = core.incorporate_detection(img, det.cuda(), [ObjectInfo(5), ObjectInfo(6)])
    assert p.shape == (3, H, W) and core.object_manager.all_obj_ids == [6, 5]  # larger segment is painted/added first
    p = core.step(img + 0.05 * torch.randn(3, H, W, generator=g).cuda())
    assert p.shape == (3, H, W) and bool(torch.isfinite(p).all())
    # a second detection frame: ids stay consistent between the object manager and the memory bank, objects that
    # keep missing detections are purged after max_missed_detection_count
    det2 = torch.zeros(H, W, dtype=torch.long)
    det2[8:40, 8:60] = 77
    for _ in range(3):
        p = core.incorporate_detection(img, det2.cuda(), [ObjectInfo(77)])
        ids = core.object_manager.all_obj_ids
        assert sorted(o for objs in core.memory.work_mem.buckets.values() for o in objs) == sorted(ids)
        assert p.shape == (len(ids) + 1, H, W) and bool(torch.isfinite(p).all())
        assert all(o.poke_count <= 1 for o in core.object_manager.obj_to_tmp_id)
    p = core.step(img)
    assert p.shape[0] == core.object_manager.num_obj + 1 and bool(torch.isfinite(p).all())
