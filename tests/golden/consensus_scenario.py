"""Seeded scenario for the in-clip consensus fixtures, shared by make_golden.py (reference side), the oracle tests and
the GPU tests.  Pure torch - imports nothing from the reference, the product or the oracle."""
import torch

H, W = 90, 120  # padded to 96 x 128 by the consensus code (6 x 8 = 48 key positions >= top_k)
TIMES = [10, 11, 12, 13]

# Objects drift by (+2, -3) pixels per frame, which `shifted_alignment` undoes exactly; jitter in the boxes sets the IoUs.
# per frame: list of (segment id in that frame's id map, (y0, y1, x0, x1), category_id, isthing, score); later boxes
# overwrite earlier ones in the id map.
#   A: seen in frames 0, 1, 3 - mutually IoU > 0.5 (a triangle of conflicts: one representative survives)
#   B: frames 0, 1, 3 with IoU(0,1), IoU(1,3) > 0.5 > IoU(0,3): a chain - the middle detection has the most support
#   C, D: seen once (no support -> dropped);  S: "stuff" on top of A (never matched with a "thing")
#   N: isthing=None in frames 1 and 3 (matched among themselves; an exact tie -> lowest index)
DETECTIONS = [
    [(3, (10, 50, 40, 90), 1, True, 0.9), (5, (56, 86, 40, 80), 2, True, 0.8)],
    [(1, (13, 52, 37, 88), 1, True, 0.7), (2, (4, 30, 90, 118), 3, True, 0.6), (4, (58, 88, 43, 83), 2, True, 0.95),
     (6, (70, 88, 96, 116), None, None, None)],
    [],
    [(9, (18, 56, 32, 80), 4, False, 0.3), (7, (16, 58, 30, 81), 1, True, 0.5), (8, (62, 92, 45, 85), 2, True, 0.4),
     (2, (2, 14, 2, 30), 5, True, 0.2), (6, (74, 90, 90, 110), None, None, 0.1)],
]


def frames(seed: int = 21):
    """[(image [3,H,W] float32, id map [H,W] int64)] - temporally correlated noise, rectangles as detections.
    Later rectangles overwrite earlier ones where they overlap (an id map holds one id per pixel)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(3, H, W, generator=g)
    out = []
    for dets in DETECTIONS:
        image = base + 0.2 * torch.randn(3, H, W, generator=g)
        ids = torch.zeros(H, W, dtype=torch.long)
        for sid, (y0, y1, x0, x1), *_ in dets:
            ids[y0:y1, x0:x1] = sid
        out.append((image, ids))
    return out


def shifted_alignment(src_ti, src_image, src_mask, tar_ti, tar_image, *unused):
    """Stand-in for spatial_alignment with a known answer: every source mask moves by (2, -3) pixels per frame of
    distance to the target; the background channel is the constant 0.5 (like the keyframe's own projection)."""
    d = tar_ti - src_ti
    moved = torch.roll(src_mask, shifts=(2 * d, -3 * d), dims=(1, 2))
    return torch.cat([torch.ones_like(moved[0:1]) * 0.5, moved], dim=0).unsqueeze(0)


# ---------------------------------------------------------------------------------------------------------------------
# match_and_merge scenario (segment_merging.py:17-143): tracked objects as (id, category, isthing, score) and three
# detection rounds; rectangles are (y0, y1, x0, x1) on a 40 x 60 grid.
MERGE_HW = (40, 60)
MERGE_TRACKED = [(10, 1, True, 0.9), (11, 2, True, 0.8), (12, 7, False, 0.7), (13, None, None, None)]
MERGE_OUR_BOXES = {1: (2, 18, 2, 20), 2: (20, 38, 4, 24), 3: (2, 30, 30, 58), 4: (32, 39, 40, 58)}  # by temporary id
MERGE_ROUNDS = [
    # (detections [(id, box, category, isthing, score)], incremental_mode, max_num_objects, our boxes override or None)
    ([(21, (3, 19, 3, 21), 1, True, 0.6),      # IoU > 0.5 with object 10 -> merged
      (22, (0, 8, 22, 28), 3, True, 0.5),      # new thing
      (23, (4, 30, 32, 58), 7, False, 0.4),    # stuff, IoU > 0.5 with stuff 12 -> merged
      (24, (30, 38, 14, 26), 2, True, 0.3),    # overlaps object 11 with IoU < 0.5 -> new object
      (25, (32, 39, 42, 58), None, None, 0.2)  # untyped, IoU > 0.5 with untyped 13
      ], False, -1, None),
    ([(31, (20, 38, 4, 24), 2, True, 0.9)], True, -1, {1: (2, 18, 2, 20), 2: (20, 38, 4, 24), 3: (0, 0, 0, 0)}),
    ([(41, (0, 5, 50, 60), 9, True, 0.9), (42, (2, 18, 2, 20), 1, True, 0.9)], False, 6, None),
]


def merge_masks(boxes, hw=MERGE_HW):
    m = torch.zeros(*hw, dtype=torch.long)
    for label, (y0, y1, x0, x1) in boxes.items():
        m[y0:y1, x0:x1] = label
    return m


# ---------------------------------------------------------------------------------------------------------------------
# ObjectManager script (object_manager.py:8-168): the same operations are replayed on the reference (when the fixture is
# minted) and on the product; `snapshot` is what gets compared after every operation.
def object_manager_script(ObjectManager, ObjectInfo, np_seed: int = 11):
    import numpy as np
    np.random.seed(np_seed)
    om = ObjectManager()
    log = []

    def snapshot(tag, extra=None):
        log.append({'op': tag, 'extra': extra,
                    'tmp_to_obj': [[int(t), int(o.id)] for t, o in om.tmp_id_to_obj.items()],
                    'obj_to_tmp': [[int(o.id), int(t)] for o, t in om.obj_to_tmp_id.items()],
                    'history': sorted(int(i) for i in om.all_historical_object_ids),
                    'all_obj_ids': [int(i) for i in om.all_obj_ids], 'num_obj': int(om.num_obj),
                    'segments': om.get_current_segments_info()})

    snapshot('init')
    snapshot('add ints', [list(map(int, r)) for r in om.add_new_objects([1, 2, 5])])
    snapshot('add colliding int', [list(map(int, r)) for r in om.add_new_objects([2])])  # random re-id in 1..255
    snapshot('add infos', [list(map(int, r)) for r in om.add_new_objects(
        [ObjectInfo(9, category_id=3, isthing=True, score=0.5), ObjectInfo(1, category_id=4, isthing=False, score=0.25)])])
    om.find_object_by_id(5).merge(ObjectInfo(77, category_id=8, isthing=True, score=0.75))
    om.find_object_by_id(5).merge(ObjectInfo(78, category_id=8, isthing=True, score=0.5))
    snapshot('merge meta into 5')
    mask = torch.arange(7).repeat(3, 1)
    snapshot('tmp_to_obj_cls', om.tmp_to_obj_cls(mask).tolist())
    snapshot('make_one_hot', om.make_one_hot(om.tmp_to_obj_cls(mask)).to(torch.uint8).tolist())
    om.delete_object(2)
    snapshot('delete 2')
    for oid, pokes in ((1, 3), (9, 6)):
        for _ in range(pokes):
            om.find_object_by_id(oid).poke()
    snapshot('purge > 4', [list(map(int, r)) if isinstance(r, list) else bool(r) for r in om.purge_inactive_objects(4)])
    snapshot('purge > 0', [list(map(int, r)) if isinstance(r, list) else bool(r) for r in om.purge_inactive_objects(0)])
    om.use_long_id = True
    snapshot('add with long ids', [list(map(int, r)) for r in om.add_new_objects([5, 300])])  # 5 < 256 -> re-id >= 256
    snapshot('has_all of an unknown id', bool(om.has_all([4242])))  # known ids crash in the reference
    return log


# ---------------------------------------------------------------------------------------------------------------------
# Semi-online session (inference_core.py:137-290): detections merged at t = 0, 3, 5, 6, plain propagation in between.
DETECT_HW = (90, 120)
DETECT_CONFIG_EXTRA = dict(max_missed_detection_count=1, max_num_objects=-1, mem_every=2)
# t -> None (step) or list of (id, box, category, isthing, score)
DETECT_SESSION = [
    [(3, (10, 50, 20, 70), 1, True, 0.9), (5, (55, 85, 60, 110), 2, True, 0.8)],
    None,
    None,
    [(1, (10, 50, 20, 70), 1, True, 0.7), (4, (4, 30, 84, 118), 3, True, 0.6)],
    None,
    [],
    [(8, (60, 88, 4, 40), 5, False, 0.5)],
    None,
]


def detect_frames(seed: int = 33):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(3, *DETECT_HW, generator=g)
    return [base + 0.2 * torch.randn(3, *DETECT_HW, generator=g) for _ in DETECT_SESSION]
