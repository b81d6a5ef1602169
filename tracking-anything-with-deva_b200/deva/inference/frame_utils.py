"""Buffered-frame record used by the semi-online detector hooks (reference deva/inference/frame_utils.py)."""
from typing import Dict, List

import torch

from deva.inference.object_info import ObjectInfo


class FrameInfo:
    def __init__(self, image: torch.Tensor, mask: torch.Tensor, segments_info: List[ObjectInfo], ti: int, info: Dict):
        self.image, self.mask, self.segments_info, self.ti, self.info = image, mask, segments_info, ti, info

    def _first(self, key):
        return self.info[key][0]

    name = property(lambda self: self._first('frame'))
    shape = property(lambda self: self.info['shape'])
    save_needed = property(lambda self: self._first('save'))
    path_to_image = property(lambda self: self._first('path_to_image'))
