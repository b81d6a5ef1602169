"""Oracle: per-frame propagation state machine (test infrastructure).

Restates deva/inference/inference_core.py:55-113,200-290, object_manager.py:27-131 and
utils/tensor_utils.py:7-48 of the reference on top of ``oracle.network`` and
``oracle.memory_bank``.  Only the VOS ``step`` path is restated (first-frame / new-object
masks and plain propagation); detection merging is out of the hot path.
"""
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import network as net
from .memory_bank import MemoryOracle


def pad_to_multiple(x: torch.Tensor, d: int = 16):
    """tensor_utils.py:7-22: symmetric zero pad, odd remainder goes to bottom/right."""
    h, w = x.shape[-2:]
    nh, nw = (h + d - 1) // d * d, (w + d - 1) // d * d
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return F.pad(x, pad), pad


def crop_pad(x: torch.Tensor, pad):
    """tensor_utils.py:25-48."""
    lw, uw, lh, uh = pad
    h, w = x.shape[-2:]
    return x[..., lh:h - uh, lw:w - uw]


class ObjectTable:
    """object_manager.py:8-168 reduced to the id maps the VOS path needs."""
    def __init__(self):
        self.tmp_to_id: Dict[int, int] = {}
        self.history = set()

    def add(self, ids: List[int]) -> List[int]:
        out = []
        for i in ids:
            new_id = i
            while new_id in self.history:  # object_manager.py:40-46 (short-id branch)
                new_id = int(np.random.randint(1, 256))
            tmp = len(self.tmp_to_id) + 1
            self.tmp_to_id[tmp] = new_id
            self.history.add(new_id)
            out.append(tmp)
        return out

    @property
    def ids(self) -> List[int]:
        return [self.tmp_to_id[t] for t in sorted(self.tmp_to_id)]

    def to_object_ids(self, tmp_mask: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(tmp_mask)
        for t, i in self.tmp_to_id.items():
            out[tmp_mask == t] = i
        return out


class CoreOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], config: Dict):
        self.sd = sd
        self.cfg = config
        self.mem_every = config['mem_every']
        self.memory = MemoryOracle(config)
        self.objects = ObjectTable()
        self.ti = -1
        self.last_mem_ti = 0
        self.last_mask = None

    def _segment(self, key, selection, ms, update_sensory=True):
        ids = self.objects.ids
        ro = self.memory.read(key, selection)
        ro = torch.stack([ro[i] for i in ids], 0).unsqueeze(0)
        sensory, _, prob = net.segment(self.sd, ms, ro, self.memory.sensory_for(ids, key),
                                       self.last_mask, update_sensory=update_sensory)
        if update_sensory:
            self.memory.set_sensory(sensory, ids)
        return prob[0]

    def _add_memory(self, image, ms, prob, key, shrinkage, selection):
        ids = self.objects.ids
        value, sensory = net.encode_mask(self.sd, image, ms, self.memory.sensory_for(ids, key),
                                         prob, deep_update=True)
        self.memory.add(key, shrinkage, value, ids, selection=selection)
        self.last_mem_ti = self.ti
        self.memory.set_sensory(sensory, ids)

    def step(self, image: torch.Tensor, mask: Optional[torch.Tensor] = None,
             objects: Optional[List[int]] = None, end: bool = False) -> torch.Tensor:
        """inference_core.py:200-290 (hard masks only)."""
        self.ti += 1
        image, pad = pad_to_multiple(image, 16)
        image = image.unsqueeze(0)
        is_mem = ((self.ti - self.last_mem_ti >= self.mem_every) or mask is not None) and not end
        need_segment = mask is None or len(self.objects.tmp_to_id) > 0
        ms, feat = net.encode_image(self.sd, image)
        key, shrinkage, selection = net.transform_key(self.sd, feat)
        if need_segment:
            prob = self._segment(key, selection, ms, update_sensory=not end)
        if mask is not None:
            tmp_ids = self.objects.add(objects)
            mask, _ = pad_to_multiple(mask, 16)
            if need_segment:
                no_bg = prob[1:]
                no_bg[:, mask > 0] = 0
                fresh = [(mask == objects[i]).type_as(no_bg).unsqueeze(0) for i in range(len(tmp_ids))]
                soft = torch.cat([no_bg, *fresh], 0)
            else:
                soft = torch.stack([mask == objects[i] for i in range(len(tmp_ids))], 0)
            prob = torch.softmax(net.aggregate(soft, dim=0), dim=0)
        self.last_mask = prob[1:].unsqueeze(0)
        if is_mem:
            self._add_memory(image, ms, self.last_mask, key, shrinkage, selection)
        return crop_pad(prob, pad)
