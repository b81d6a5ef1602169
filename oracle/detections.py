"""CPU restatement of the detection-merging side of the propagation core (test infrastructure).

    ObjectBook                      deva/inference/object_manager.py:8-168 + object_info.py:7-62 (ids, poke counts, votes)
    match_and_merge                 deva/inference/segment_merging.py:17-143
    DetectionCoreOracle             deva/inference/inference_core.py:137-198 (incorporate_detection) on top of CoreOracle

Pinned by tests/golden/detections.{npz,json}, minted from the reference by tests/golden/make_golden.py.
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from oracle import network as net
from oracle.core import CoreOracle, crop_pad, pad_to_multiple


class Tracked:
    def __init__(self, id, category_id=None, isthing=None, score=None):
        self.id, self.category_ids, self.scores, self.isthing, self.poke_count = id, [category_id], [score], isthing, 0

    def merge(self, other):
        self.category_ids.extend(other.category_ids)
        self.scores.extend(other.scores)


class ObjectBook:
    """Insertion-ordered tmp id -> Tracked; tmp ids are always 1..K (object_manager.py:68-89 re-packs on deletion)."""
    def __init__(self):
        self.by_tmp: Dict[int, Tracked] = {}
        self.history = set()

    # --- the slice of the interface CoreOracle uses
    @property
    def tmp_to_id(self) -> Dict[int, int]:
        return {t: o.id for t, o in self.by_tmp.items()}

    @property
    def ids(self) -> List[int]:
        return [o.id for o in self.by_tmp.values()]

    def add(self, objects) -> List[int]:
        out = []
        for obj in objects:
            if isinstance(obj, int):
                obj = Tracked(obj)
            new_id = obj.id
            while new_id in self.history:  # object_manager.py:40-47, short-id branch
                new_id = int(np.random.randint(1, 256))
            fresh = Tracked(new_id)
            fresh.category_ids, fresh.scores, fresh.isthing = obj.category_ids, obj.scores, obj.isthing  # copy_meta_info
            tmp = len(self.by_tmp) + 1
            self.by_tmp[tmp] = fresh
            self.history.add(new_id)
            out.append(tmp)
        return out

    def delete(self, ids: List[int]) -> None:
        keep = [o for o in self.by_tmp.values() if o.id not in ids]
        self.by_tmp = {i + 1: o for i, o in enumerate(keep)}

    def purge(self, max_missed: int):
        gone = [o.id for o in self.by_tmp.values() if o.poke_count > max_missed]
        keep_tmp = [t for t, o in self.by_tmp.items() if o.poke_count <= max_missed]
        keep_ids = [o.id for o in self.by_tmp.values() if o.poke_count <= max_missed]
        if gone:
            self.delete(gone)
        return bool(gone), keep_tmp, keep_ids

    def one_hot(self, id_mask: torch.Tensor) -> torch.Tensor:
        return torch.stack([id_mask == o.id for o in self.by_tmp.values()], 0) if self.by_tmp else \
            torch.zeros((0, *id_mask.shape), dtype=torch.bool)


def match_and_merge(our_mask: torch.Tensor, new_mask: torch.Tensor, book: ObjectBook, new_infos: List[Tracked],
                    max_num_objects: int = -1, incremental: bool = False) -> torch.Tensor:
    """segment_merging.py:92-143 + merge_by_iou :26-89.  our_mask: tmp ids; new_mask: detection ids."""
    our_mask, new_mask = our_mask.long(), new_mask.long()
    ours = {t: our_mask == t for t in book.by_tmp}
    if max_num_objects is not None and max_num_objects > 0 and len(book.by_tmp) + len(new_infos) > max_num_objects:
        new_infos = []
    news = {d.id: new_mask == d.id for d in new_infos}
    our_sum = {t: int(m.sum()) for t, m in ours.items()}
    new_sum = {i: int(m.sum()) for i, m in news.items()}
    merged = torch.zeros_like(our_mask)
    for status in (None, False, True):
        matching, area = {}, {}
        for d in new_infos:
            if d.isthing != status:
                continue
            for t, o in list(book.by_tmp.items()):
                if o.isthing != status or t in matching or t not in ours:
                    continue
                inter = int((news[d.id] & ours[t]).sum())
                if inter < 1e-3:
                    continue
                union = new_sum[d.id] + our_sum[t] - inter
                if inter / union > 0.5:
                    matching[t] = d
                    area[('our', t)] = union
                    break
            else:
                area[('new', d.id)] = new_sum[d.id]
        for t, o in book.by_tmp.items():
            if o.isthing == status and t not in matching and t in ours:
                area[('our', t)] = our_sum[t]
        by_id = {d.id: d for d in new_infos}
        snapshot = dict(book.by_tmp)  # objects added while painting must not be visited as 'ours'
        for (kind, key), _ in sorted(area.items(), key=lambda kv: kv[1], reverse=True):
            if kind == 'new':
                tmp = book.add([by_id[key]])[0]
                merged[news[key]] = book.by_tmp[tmp].id
            else:
                o = snapshot[key]
                merged[ours[key]] = o.id
                if key in matching:
                    merged[news[matching[key].id]] = o.id
                    o.merge(matching[key])
                    o.poke_count = 0
                elif incremental:
                    o.poke_count = o.poke_count + 1 if our_sum[key] < 1 else 0
                else:
                    o.poke_count += 1
    return book.one_hot(merged)


class DetectionCoreOracle(CoreOracle):
    def __init__(self, sd, config: Dict):
        super().__init__(sd, config)
        self.objects = ObjectBook()
        self.max_missed = config.get('max_missed_detection_count')
        self.max_num_objects = config.get('max_num_objects')

    def incorporate_detection(self, image: torch.Tensor, new_mask: torch.Tensor, infos: List[Tracked],
                              forward_mask: Optional[torch.Tensor] = None, incremental: bool = False) -> torch.Tensor:
        self.ti += 1
        image, pad = pad_to_multiple(image, 16)
        new_mask, _ = pad_to_multiple(new_mask, 16)
        image = image.unsqueeze(0)
        ms, feat = net.encode_image(self.sd, image)
        key, shrinkage, selection = net.transform_key(self.sd, feat)
        if forward_mask is None:
            if self.memory.engaged:
                forward_mask = torch.argmax(self._segment(key, selection, ms), dim=0)
            else:
                forward_mask = torch.zeros_like(new_mask)
        merged = match_and_merge(forward_mask, new_mask, self.objects, infos, self.max_num_objects, incremental)
        purged, keep_tmp, keep_ids = self.objects.purge(self.max_missed)
        if purged:
            self.memory.keep_only(keep_ids)
            merged = merged[[t - 1 for t in keep_tmp]]
        self.last_mask = merged.unsqueeze(0).float()
        self._add_memory(image, ms, self.last_mask, key, shrinkage, selection)
        return crop_pad(net.aggregate(self.last_mask[0], dim=0), pad)
