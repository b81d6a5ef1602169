"""Object id bookkeeping (API and semantics of the reference's deva/inference/object_manager.py:8-168).

Real object ids never change; temporary ids are the 1-based channel positions of the objects in
every per-object tensor and are re-packed when objects are deleted.  All of this is integer logic on
the host and is bit-exact with the reference (including the numpy RNG draw on id collisions).
"""
from typing import Dict, List, Set, Tuple, Union

import numpy as np
import torch

from deva.inference.object_info import ObjectInfo


class ObjectManager:
    def __init__(self):
        self.obj_to_tmp_id: Dict[ObjectInfo, int] = {}
        self.tmp_id_to_obj: Dict[int, ObjectInfo] = {}
        self.obj_id_to_obj: Dict[int, ObjectInfo] = {}
        self.all_historical_object_ids: Set[int] = set()  # ids are never reused (quirk Q2)
        self.use_long_id = False

    def _reindex(self) -> None:
        self.obj_id_to_obj = {o.id: o for o in self.obj_to_tmp_id}

    def _fresh_id(self, wanted: int) -> int:
        """The wanted id if free, else random draws exactly as object_manager.py:38-53."""
        candidate, tries = wanted, 0
        while candidate in self.all_historical_object_ids or (self.use_long_id and candidate < 256):
            candidate = np.random.randint(256, 256**3) if self.use_long_id else np.random.randint(1, 256)
            tries += 1
            if tries > 5000:
                raise ValueError('We cannot find a new ID for this object. Perhaps you should use long ID?')
        return candidate

    def add_new_objects(self, objects: Union[List[ObjectInfo], ObjectInfo, List[int]]) -> Tuple[List[int], List[int]]:
        if not isinstance(objects, list):
            objects = [objects]
        tmp_ids, obj_ids = [], []
        for src in objects:
            if isinstance(src, int):
                src = ObjectInfo(id=src)
            obj = ObjectInfo(id=self._fresh_id(src.id))
            obj.copy_meta_info(src)
            tmp = len(self.obj_to_tmp_id) + 1
            self.obj_to_tmp_id[obj] = tmp
            self.tmp_id_to_obj[tmp] = obj
            self.all_historical_object_ids.add(obj.id)
            tmp_ids.append(tmp)
            obj_ids.append(obj.id)
        self._reindex()
        assert tmp_ids == sorted(tmp_ids), 'tmp id assignment bugged'
        return tmp_ids, obj_ids

    def delete_object(self, obj_ids_to_remove: Union[int, List[int]]) -> None:
        if isinstance(obj_ids_to_remove, int):
            obj_ids_to_remove = [obj_ids_to_remove]
        survivors = [self.tmp_id_to_obj[t] for t in range(1, len(self.obj_to_tmp_id) + 1)
                     if self.tmp_id_to_obj[t].id not in obj_ids_to_remove]
        self.obj_to_tmp_id = {o: i + 1 for i, o in enumerate(survivors)}
        self.tmp_id_to_obj = {i + 1: o for i, o in enumerate(survivors)}
        self._reindex()

    def purge_inactive_objects(self, max_missed_detection_count: int) -> Tuple[bool, List[int], List[int]]:
        dead = [o for o in self.obj_to_tmp_id if o.poke_count > max_missed_detection_count]
        alive = [o for o in self.obj_to_tmp_id if o.poke_count <= max_missed_detection_count]
        tmp_keep = [self.obj_to_tmp_id[o] for o in alive]
        obj_keep = [o.id for o in alive]
        if dead:
            self.delete_object([o.id for o in dead])
        return len(dead) > 0, tmp_keep, obj_keep

    def tmp_to_obj_cls(self, mask: torch.Tensor) -> torch.Tensor:
        """Class map in temporary ids -> class map in real object ids, via one table lookup."""
        k = len(self.tmp_id_to_obj)
        table = torch.zeros(k + 1, dtype=mask.dtype)
        for tmp, obj in self.tmp_id_to_obj.items():
            table[tmp] = obj.id
        table = table.to(mask.device)
        m = mask.long()
        known = (m >= 0) & (m <= k)
        return torch.where(known, table[m.clamp(0, k)], torch.zeros_like(mask))

    def get_tmp_to_obj_mapping(self) -> Dict[int, int]:
        # {object id: tmp id}, picklable.  The reference's version (object_manager.py:119-121) swaps its
        # loop variables and raises on any call; this is its documented intent.
        return {obj.id: tmp for tmp, obj in self.tmp_id_to_obj.items()}

    def realize_dict(self, obj_dict: Dict[int, torch.Tensor]) -> torch.Tensor:
        """Dict keyed by object id -> tensor stacked in temporary-id order.

        When the per-object tensors are already consecutive slices of one buffer in that order
        (what MemoryManager.match_memory produces) the buffer is returned without a copy.
        """
        parts = []
        for _, obj in self.tmp_id_to_obj.items():
            if obj.id not in obj_dict:
                raise NotImplementedError
            parts.append(obj_dict[obj.id])
        first = parts[0]
        if len(parts) > 1 and all(p.shape == first.shape and p.stride() == first.stride() and p.dtype == first.dtype and
                                  p.untyped_storage().data_ptr() == first.untyped_storage().data_ptr() for p in parts):
            step = parts[1].storage_offset() - first.storage_offset()
            if step > 0 and all(p.storage_offset() == first.storage_offset() + i * step for i, p in enumerate(parts)):
                return first.as_strided((len(parts), *first.shape), (step, *first.stride()), first.storage_offset())
        return torch.stack(parts, dim=0)

    def make_one_hot(self, cls_mask: torch.Tensor) -> torch.Tensor:
        ids = [obj.id for _, obj in self.tmp_id_to_obj.items()]
        if not ids:
            return torch.zeros((0, *cls_mask.shape), dtype=torch.bool, device=cls_mask.device)
        table = torch.tensor(ids, dtype=cls_mask.dtype, device=cls_mask.device)
        return cls_mask.unsqueeze(0) == table.view(-1, *([1] * cls_mask.dim()))

    def get_current_segments_info(self) -> List[Dict]:
        return [{'category_id': o.vote_category_id(), 'id': int(o.id), 'score': o.vote_score()}
                for o in self.obj_to_tmp_id]

    @property
    def all_obj_ids(self) -> List[int]:
        return [o.id for o in self.obj_to_tmp_id]

    @property
    def num_obj(self) -> int:
        return len(self.obj_to_tmp_id)

    def has_all(self, objects: List[int]) -> bool:
        """True when every id is already tracked.

        The reference raises AttributeError here for a tracked id (``int in Dict[ObjectInfo]``,
        SURVEY quirk Q1); this is the evident intent, and agrees with it whenever it returns.
        """
        return all(i in self.obj_id_to_obj for i in objects)

    def find_object_by_id(self, obj_id) -> ObjectInfo:
        return self.obj_id_to_obj[obj_id]
