"""Frame ingest / egress around ``DEVAInferenceCore.step``: the two per-frame host<->device hops of a
real pipeline, each as one kernel.

* ``frame_from_rgb8``: the decoded uint8 frame is uploaded as is (3 bytes / pixel instead of 12) and ToTensor +
  Normalize (deva/inference/data/video_reader.py:146-150, IMAGENET mean/std) run on the device - bit-exact.
* ``prob_to_ids``: the driver's post-step (evaluation/eval_vos.py:169-181) - bilinear resize to the original size,
  optional flip, argmax, ``ObjectManager.tmp_to_obj_cls`` - fused into a single pass that writes the id map.
"""
from typing import Optional, Sequence, Tuple

import torch

from deva import _native as nat

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def frame_from_rgb8(frame: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                    device: Optional[torch.device] = None) -> torch.Tensor:
    """uint8 [H, W, 3] RGB (host - ideally pinned - or device) -> normalised float32 [3, H, W] on the device."""
    assert frame.dtype == torch.uint8 and frame.dim() == 3 and frame.shape[2] == 3
    if not frame.is_cuda:
        frame = frame.to(device or torch.device('cuda', torch.cuda.current_device()), non_blocking=True)
    frame = frame.contiguous()
    h, w = frame.shape[:2]
    out = torch.empty(3, h, w, dtype=torch.float32, device=frame.device)
    nat.ingest_rgb8(frame, out, h, w, mean, std)
    return out


def id_lut(object_manager, channels: int, device) -> torch.Tensor:
    """int32 [channels]: temporary id (prob channel) -> object id; channel 0 and unknown channels -> 0."""
    lut = [0] * channels
    for tmp_id, obj in object_manager.tmp_id_to_obj.items():
        if 0 < tmp_id < channels:
            lut[tmp_id] = int(obj.id)
    return torch.tensor(lut, dtype=torch.int32, device=device)


def prob_to_ids(prob: torch.Tensor, object_manager=None, size: Optional[Tuple[int, int]] = None, flip: bool = False,
                dtype: torch.dtype = torch.long, lut: Optional[torch.Tensor] = None) -> torch.Tensor:
    """prob float32 [K+1, H, W] from ``step`` -> id map [H0, W0] (``size`` or H, W), dtype long or uint8.

    Equals ``tmp_to_obj_cls(argmax(flip(interpolate(prob, size, 'bilinear', align_corners=False))))``."""
    assert prob.is_cuda and prob.dtype == torch.float32 and prob.dim() == 3 and dtype in (torch.long, torch.uint8)
    prob = prob.contiguous()
    c, h, w = prob.shape
    oh, ow = (h, w) if size is None else (int(size[0]), int(size[1]))
    if lut is None and object_manager is not None:
        lut = id_lut(object_manager, c, prob.device)
    out = torch.empty(oh, ow, dtype=dtype, device=prob.device)
    nat.prob_to_ids(prob, c, h, w, oh, ow, flip, lut, out if dtype == torch.uint8 else None,
                    out if dtype == torch.long else None)
    return out
