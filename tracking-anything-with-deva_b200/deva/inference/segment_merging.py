"""Merging propagated segments with image-level detections (paper section 3.2.2).

Same contract and decisions as the reference's deva/inference/segment_merging.py:17-143 (greedy IoU > 0.5
matching per thing/stuff class in detection order, area-ordered painting, poke/unpoke bookkeeping),
but every pixel count comes from ONE joint label histogram (one device reduction + one host transfer)
instead of a ``.sum()`` per mask pair.  Integer logic; identical results.
"""
import warnings
from typing import Dict, List, Literal, Tuple

import torch

from deva.inference.object_info import ObjectInfo
from deva.inference.object_manager import ObjectManager


def _joint_histogram(our_mask: torch.Tensor, new_slot: torch.Tensor, n_our: int, n_new: int) -> List[List[int]]:
    """counts[t][s] = #pixels with temporary id t (0..n_our) and detection slot s (0..n_new)."""
    joint = our_mask.reshape(-1).clamp(0, n_our) * (n_new + 1) + new_slot.reshape(-1)
    counts = torch.bincount(joint, minlength=(n_our + 1) * (n_new + 1))
    return counts.view(n_our + 1, n_new + 1).tolist()


def match_and_merge(our_mask: torch.Tensor, new_mask: torch.Tensor, object_manager: ObjectManager,
                    new_segments_info: List[ObjectInfo], mode: Literal['iou'] = 'iou', max_num_objects: int = -1,
                    incremental_mode: bool = False) -> torch.Tensor:
    """our_mask: temporary ids [H,W]; new_mask: detection ids [H,W].  Returns one-hot [K,H,W] (bool) in the
    temporary-id order of the updated object manager; updates the object manager as a side effect."""
    if mode.lower() != 'iou':
        raise NotImplementedError('Engulf mode is deprecated')
    our_mask, new_mask = our_mask.long(), new_mask.long()
    if max_num_objects is not None and max_num_objects > 0 and \
            len(object_manager.obj_to_tmp_id) + len(new_segments_info) > max_num_objects:
        warnings.warn('Number of objects exceeded maximum (--max_num_objects); discarding new objects')
        new_segments_info = []

    ours: List[Tuple[ObjectInfo, int]] = list(object_manager.obj_to_tmp_id.items())  # insertion order
    n_our = len(ours)
    # detection slot s (1-based) for every listed detection id; pixels of unlisted ids fall in slot 0
    slot_of: Dict[int, int] = {}
    for det in new_segments_info:
        slot_of.setdefault(det.id, len(slot_of) + 1)
    n_new = len(slot_of)
    new_slot = torch.zeros_like(new_mask)
    for det_id, s in slot_of.items():
        new_slot[new_mask == det_id] = s
    counts = _joint_histogram(our_mask, new_slot, n_our, n_new)
    our_area = {obj: sum(counts[tmp]) for obj, tmp in ours}
    new_area = {det: sum(counts[t][slot_of[det.id]] for t in range(n_our + 1)) for det in new_segments_info}

    merged = torch.zeros_like(our_mask)
    for isthing in (None, False, True):  # stuff / things / unlabelled are merged separately
        matched: Dict[ObjectInfo, ObjectInfo] = {}
        area: Dict[Tuple[ObjectInfo, bool], int] = {}
        for det in new_segments_info:
            if det.isthing != isthing:
                continue
            for obj, tmp in ours:
                if obj.isthing != isthing or obj in matched:
                    continue
                inter = counts[tmp][slot_of[det.id]]
                if inter < 1e-3:
                    continue
                union = new_area[det] + our_area[obj] - inter
                if inter / union > 0.5:
                    matched[obj] = det
                    area[(obj, False)] = union
                    break
            else:
                area[(det, True)] = new_area[det]
        for obj, _ in ours:
            if obj.isthing == isthing and obj not in matched:
                area[(obj, False)] = our_area[obj]

        # paint large segments first so small ones stay visible
        for (obj, is_new), _ in sorted(area.items(), key=lambda kv: kv[1], reverse=True):
            if is_new:
                _, new_ids = object_manager.add_new_objects(obj)
                merged[new_slot == slot_of[obj.id]] = new_ids[0]
                continue
            tmp = dict(ours)[obj]
            merged[our_mask == tmp] = obj.id
            if obj in matched:
                det = matched[obj]
                merged[new_slot == slot_of[det.id]] = obj.id
                obj.merge(det)
                obj.unpoke()
            elif incremental_mode and our_area[obj] >= 1:
                obj.unpoke()
            else:
                obj.poke()
    return object_manager.make_one_hot(merged)
