"""CPU tests: host-side id/padding/merging logic and the C-ABI export surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle.core import ObjectTable, crop_pad, pad_to_multiple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pad_unpad_match_oracle():
    from deva.utils.tensor_utils import pad_divide_by, unpad
    for h, w in [(480, 854), (1080, 1920), (96, 96), (17, 33), (5, 7)]:
        x = torch.arange(3 * h * w, dtype=torch.float32).view(3, h, w)
        y, pad = pad_divide_by(x, 16)
        ref, rpad = pad_to_multiple(x, 16)
        assert tuple(pad) == tuple(rpad) and torch.equal(y, ref)
        assert y.shape[-1] % 16 == 0 and y.shape[-2] % 16 == 0
        assert torch.equal(unpad(y, pad), x) and torch.equal(crop_pad(ref, rpad), x)
        assert torch.equal(unpad(y.unsqueeze(0), pad), x.unsqueeze(0))


def test_object_manager_ids_bit_exact():
    from deva.inference.object_manager import ObjectManager
    np.random.seed(7)
    om = ObjectManager()
    tmp, ids = om.add_new_objects([1, 2, 5])
    assert tmp == [1, 2, 3] and ids == [1, 2, 5]
    tmp2, ids2 = om.add_new_objects([2])  # collision -> random re-id drawn from numpy's global RNG
    np.random.seed(7)
    ref = ObjectTable()
    ref.add([1, 2, 5])
    ref.add([2])
    assert {t: o.id for t, o in om.tmp_id_to_obj.items()} == ref.tmp_to_id
    mask = torch.tensor([[0, 1, 2], [3, 4, 4]])
    assert torch.equal(om.tmp_to_obj_cls(mask), ref.to_object_ids(mask))
    assert om.all_obj_ids == ref.ids and om.num_obj == 4
    assert om.has_all([1, 5]) and not om.has_all([1, 99])
    om.delete_object([2])
    assert [o.id for o in om.tmp_id_to_obj.values()] == [1, 5, ids2[0]]
    assert list(om.tmp_id_to_obj.keys()) == [1, 2, 3]
    one_hot = om.make_one_hot(torch.tensor([[1, 5], [0, ids2[0]]]))
    assert one_hot.shape == (3, 2, 2) and one_hot.sum().item() == 3
    d = {o.id: torch.full((2, 2), float(o.id)) for o in om.tmp_id_to_obj.values()}
    assert om.realize_dict(d)[:, 0, 0].tolist() == [1.0, 5.0, float(ids2[0])]
    # zero-copy path: consecutive slices of one buffer come back as a view
    buf = torch.arange(12.).view(3, 2, 2)
    view = om.realize_dict({o.id: buf[i] for i, o in enumerate(om.tmp_id_to_obj.values())})
    assert view.data_ptr() == buf.data_ptr() and torch.equal(view, buf)


def test_purge_inactive():
    from deva.inference.object_manager import ObjectManager
    om = ObjectManager()
    om.add_new_objects([3, 4, 9])
    om.find_object_by_id(4).poke_count = 6
    purged, tmp_keep, obj_keep = om.purge_inactive_objects(5)
    assert purged and tmp_keep == [1, 3] and obj_keep == [3, 9]
    assert {t: o.id for t, o in om.tmp_id_to_obj.items()} == {1: 3, 2: 9}


def _merge_reference(our_mask, new_mask, our, dets):
    """Straight restatement of the reference's greedy IoU merge (segment_merging.py:17-143), ids only."""
    merged = torch.zeros_like(our_mask)
    next_tmp = len(our)
    result_ids = dict(our)  # obj id -> tmp
    area, matched = {}, {}
    for d in dets:
        dm = new_mask == d
        for oid, tmp in our.items():
            if oid in matched:
                continue
            om_ = our_mask == tmp
            inter = int((dm & om_).sum())
            if inter == 0:
                continue
            union = int(dm.sum()) + int(om_.sum()) - inter
            if inter / union > 0.5:
                matched[oid] = d
                area[(oid, False)] = union
                break
        else:
            area[(d, True)] = int(dm.sum())
    for oid, tmp in our.items():
        if oid not in matched:
            area[(oid, False)] = int((our_mask == tmp).sum())
    for (x, is_new), _ in sorted(area.items(), key=lambda kv: kv[1], reverse=True):
        if is_new:
            next_tmp += 1
            result_ids[x] = next_tmp
            merged[new_mask == x] = x
        else:
            merged[our_mask == our[x]] = x
            if x in matched:
                merged[new_mask == matched[x]] = x
    return merged, result_ids


def test_match_and_merge_matches_reference_logic():
    from deva.inference.object_info import ObjectInfo
    from deva.inference.object_manager import ObjectManager
    from deva.inference.segment_merging import match_and_merge
    g = torch.Generator().manual_seed(0)
    for trial in range(5):
        H = W = 24
        our_mask = torch.zeros(H, W, dtype=torch.long)
        our_mask[2:12, 2:12] = 1
        our_mask[12:22, 4:20] = 2
        new_mask = torch.zeros(H, W, dtype=torch.long)
        r = int(torch.randint(0, 4, (1, ), generator=g))
        new_mask[2 + r:12 + r, 2:12] = 40          # overlaps object tmp 1
        new_mask[0:6, 14:24] = 41                  # new object
        new_mask[18:24, 0:3] = 42                  # new object, small
        om = ObjectManager()
        om.add_new_objects([10, 11])
        dets = [ObjectInfo(40), ObjectInfo(41), ObjectInfo(42)]
        out = match_and_merge(our_mask, new_mask, om, dets)
        want, ids = _merge_reference(our_mask, new_mask, {10: 1, 11: 2}, [40, 41, 42])
        got_cls = torch.zeros_like(our_mask)
        for tmp, obj in om.tmp_id_to_obj.items():
            got_cls[out[tmp - 1]] = obj.id
        assert torch.equal(got_cls, want), trial
        assert {o.id: t for o, t in om.obj_to_tmp_id.items()} == ids
        assert om.find_object_by_id(11).poke_count == 1


def test_match_and_merge_matches_reference_fixture(golden_dir):
    """Three detection rounds (plain, incremental, object cap) against outputs of the reference's match_and_merge
    (tests/golden/make_golden.py::golden_match_and_merge): one-hot masks, ids, poke counts and merged meta bit-exact."""
    import importlib.util
    import json
    import warnings
    from deva.inference.object_info import ObjectInfo
    from deva.inference.object_manager import ObjectManager
    from deva.inference.segment_merging import match_and_merge
    spec = importlib.util.spec_from_file_location('consensus_scenario', os.path.join(golden_dir, 'consensus_scenario.py'))
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    g = np.load(os.path.join(golden_dir, 'match_and_merge.npz'))
    want = json.load(open(os.path.join(golden_dir, 'match_and_merge.json')))['rounds']
    np.random.seed(5)
    om = ObjectManager()
    om.add_new_objects([ObjectInfo(i, category_id=c, isthing=t, score=s) for i, c, t, s in sc.MERGE_TRACKED])
    our_boxes = dict(sc.MERGE_OUR_BOXES)
    for r, (dets, incremental, cap, override) in enumerate(sc.MERGE_ROUNDS):
        if override is not None:
            our_boxes = dict(override)
        infos = [ObjectInfo(d[0], category_id=d[2], isthing=d[3], score=d[4]) for d in dets]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            one_hot = match_and_merge(sc.merge_masks(our_boxes), sc.merge_masks({d[0]: d[1] for d in dets}), om, infos,
                                      max_num_objects=cap, incremental_mode=incremental)
        assert torch.equal(one_hot.to(torch.uint8), torch.from_numpy(g[f'merge_{r}'])), r
        state = [[t, o.id, o.poke_count, o.category_ids, o.scores] for t, o in om.tmp_id_to_obj.items()]
        assert state == want[r], (r, state, want[r])


def test_object_manager_script_matches_reference_fixture(golden_dir):
    """Ids, random re-ids (numpy global RNG), deletion with re-packing, purging, votes, id-map conversions: every
    snapshot of the scripted session equals the one recorded from the reference's ObjectManager."""
    import importlib.util
    import json
    from deva.inference.object_info import ObjectInfo
    from deva.inference.object_manager import ObjectManager
    spec = importlib.util.spec_from_file_location('consensus_scenario', os.path.join(golden_dir, 'consensus_scenario.py'))
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    want = json.load(open(os.path.join(golden_dir, 'object_manager.json')))
    got = json.loads(json.dumps(sc.object_manager_script(ObjectManager, ObjectInfo)))
    assert len(got) == len(want)
    for a_, b_ in zip(got, want):
        assert a_ == b_, (a_['op'], a_, b_)


def test_eval_args_match_reference_fixture(golden_dir):
    """add_common_eval_args (eval_args.py:7-44): same flags, defaults, types and store_true actions as the reference
    (fixture = the reference parser's actions, dumped when the fixtures were minted)."""
    import json
    from argparse import ArgumentParser
    from deva.inference.eval_args import add_common_eval_args
    p = ArgumentParser()
    add_common_eval_args(p)
    mine = {a.dest: [a.default, type(a).__name__, (a.type.__name__ if a.type else None)] for a in p._actions if a.dest != 'help'}
    assert mine == json.load(open(os.path.join(golden_dir, 'eval_args.json')))


def test_pad_unpad_roundtrip_property():
    """pad_divide_by / unpad (tensor_utils.py:7-48) for arbitrary sizes: multiple of d, symmetric with the odd pixel at
    the bottom/right, zero filled, exact round trip - and identical to the oracle."""
    from hypothesis import given, settings, strategies as st
    from deva.utils.tensor_utils import pad_divide_by, unpad

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 70), st.integers(1, 70), st.sampled_from([8, 16, 32]))
    def check(h, w, d):
        x = torch.arange(2 * h * w, dtype=torch.float32).view(2, h, w) + 1
        y, pad = pad_divide_by(x, d)
        assert y.shape[-2] % d == 0 and y.shape[-1] % d == 0 and y.shape[-2] - h < d and y.shape[-1] - w < d
        lw, uw, lh, uh = pad
        assert 0 <= uw - lw <= 1 and 0 <= uh - lh <= 1
        assert float(y.double().sum()) == float(x.double().sum())  # padding is zeros
        assert torch.equal(unpad(y, pad), x)
        ref, rpad = pad_to_multiple(x, d)
        assert tuple(pad) == tuple(rpad) and torch.equal(y, ref)

    check()


def test_id_lut_matches_tmp_to_obj_cls():
    """frame_io.id_lut is the table form of ObjectManager.tmp_to_obj_cls (object_manager.py:112-117)."""
    from deva.inference.frame_io import id_lut
    from deva.inference.object_manager import ObjectManager
    np.random.seed(3)
    om = ObjectManager()
    om.add_new_objects([4, 9, 250, 7])
    om.delete_object([9])
    lut = id_lut(om, 6, 'cpu')
    mask = torch.randint(0, 4, (5, 7))
    assert lut.dtype == torch.int32 and lut[0] == 0
    assert torch.equal(lut[mask].long(), om.tmp_to_obj_cls(mask))
    assert lut.tolist()[len(om.tmp_id_to_obj) + 1:] == [0] * (6 - len(om.tmp_id_to_obj) - 1)


def test_package_chains_to_reference_checkout(tmp_path):
    """SURVEY 8(b) "imports that must resolve": with this package first on sys.path and a reference checkout after it,
    modules we provide load from here, everything else (dataset readers, deva.ext, ...) from the checkout."""
    import subprocess
    import sys
    import textwrap
    for d in ('deva', 'deva/inference', 'deva/inference/data', 'deva/ext', 'deva/utils'):
        (tmp_path / d).mkdir(parents=True, exist_ok=True)
        (tmp_path / d / '__init__.py').write_text('')
    (tmp_path / 'deva/inference/data/fake_reader.py').write_text(
        'from deva.inference.object_info import ObjectInfo\nKIND = "reference reader"\n')
    (tmp_path / 'deva/inference/object_info.py').write_text('raise RuntimeError("shadowed module must not load")\n')
    (tmp_path / 'deva/utils/pano_utils.py').write_text('def id_to_rgb(i): return (i, 0, 0)\n')
    pkg = os.path.join(ROOT, 'tracking-anything-with-deva_b200')
    code = textwrap.dedent("""
        import deva.inference.data.fake_reader as r, deva.inference.object_info as oi, deva.utils.pano_utils as pu, deva.ext
        print(r.KIND, '|', oi.__file__, '|', pu.id_to_rgb(3), '|', r.ObjectInfo is oi.ObjectInfo)
    """)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True,
                         env=dict(os.environ, PYTHONPATH=pkg + os.pathsep + str(tmp_path)))
    assert out.returncode == 0, out.stderr
    kind, oi_file, rgb, same = [t.strip() for t in out.stdout.strip().split('|')]
    assert kind == 'reference reader' and oi_file.startswith(pkg) and rgb == '(3, 0, 0)' and same == 'True'


def test_consensus_selection_is_optimal():
    """solve_exact (no ILP solver needed) == exhaustive enumeration of the reference's integer program
    (consensus_automatic.py:28-79) on random conflict graphs, including the tie-breaking rule."""
    import numpy as np
    from deva.inference.consensus_automatic import solve_exact
    from oracle.consensus import solve_brute_force
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(1, 12))
        iou = np.zeros((n, n), dtype=np.float32)
        for i in range(n):
            for j in range(i + 1, n):
                if rng.random() < 0.25:
                    iou[i, j] = iou[j, i] = np.float32(rng.choice([0.55, 0.6, 0.75, 0.9]) if trial % 2 else rng.uniform(0.5, 1))
        ind = iou > 0.49
        assert solve_exact(iou * ind, ind, n) == solve_brute_force(iou * ind, ind, n), (trial, iou)


def test_c_abi_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, 'tracking-anything-with-deva_b200', 'csrc', 'libdeva_b200.so')
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    header = open(os.path.join(ROOT, 'include', 'deva_b200.h')).read()
    declared = set(re.findall(r'DEVA_B200_API[^;(]*?\b(deva_b200_\w+)\s*\(', header))
    assert len(declared) >= 32
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    from deva import _native
    assert set(_native.EXPORTS) == declared
    assert _native.lib().deva_b200_abi_version() == _native.ABI_VERSION  # loads without a GPU; no compute is launched


def test_native_refuses_to_run_without_gpu():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from deva import _native
    with pytest.raises(RuntimeError):
        _native.require_device()
    from deva.model.network import DEVA
    net = DEVA(dict(key_dim=64, value_dim=512, pix_feat_dim=512))
    with pytest.raises(RuntimeError):
        net.encode_image(torch.zeros(1, 3, 32, 32))
