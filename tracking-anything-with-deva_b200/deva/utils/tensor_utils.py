"""Padding helpers at the frame boundary (API of the reference's deva/utils/tensor_utils.py:7-48)."""
from typing import Iterable, Tuple

import torch
import torch.nn.functional as F


def pad_divide_by(in_img: torch.Tensor, d: int) -> Tuple[torch.Tensor, Tuple[int, int, int, int]]:
    """Zero-pad the last two dims up to multiples of ``d``; an odd remainder goes to the bottom/right.

    Returns (padded, (left, right, top, bottom)) like the reference (tensor_utils.py:7-22).
    """
    h, w = in_img.shape[-2:]
    extra_h, extra_w = (-h) % d, (-w) % d
    top, left = extra_h // 2, extra_w // 2
    pad = (left, extra_w - left, top, extra_h - top)
    if extra_h == 0 and extra_w == 0:
        return in_img, pad
    return F.pad(in_img, pad), pad


def unpad(img: torch.Tensor, pad: Iterable[int]) -> torch.Tensor:
    """Inverse of pad_divide_by for 2-D..5-D tensors (tensor_utils.py:25-48)."""
    left, right, top, bottom = pad
    if img.dim() < 2 or img.dim() > 5:
        raise NotImplementedError
    h, w = img.shape[-2:]
    return img[..., top:h - bottom, left:w - right]
