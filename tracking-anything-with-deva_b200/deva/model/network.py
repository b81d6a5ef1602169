"""``DEVA`` - the network object the inference core drives (API of deva/model/network.py:18-190).

Holds the checkpoint tensors under the reference's own names (so ``state_dict()`` /
``load_state_dict()`` / ``load_weights()`` exchange checkpoints with the reference unchanged) and
runs inference through ``deva.model.engine.Engine``.  Inference only: the training entry points of the
reference (``read_memory``, ``need_aux`` heads, ``forward(mode, ...)`` dispatch used by DDP) are not part
of the propagation hot path.
"""
from typing import Dict, Iterable, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from deva.model.engine import Engine
from deva.model.native_engine import NativeEngine
from deva.model.param_spec import checkpoint_spec, synthetic_state_dict


class _Node(nn.Module):
    """Anonymous container used to reproduce the dotted checkpoint names."""


class DEVA(nn.Module):
    def __init__(self, config: Dict, *, init_seed: int = 0):
        super().__init__()
        self.pix_feat_dim = config['pix_feat_dim']
        self.key_dim = config['key_dim']
        self.value_dim = config['value_dim']
        init = synthetic_state_dict(init_seed, self.key_dim, self.value_dim, self.pix_feat_dim)
        for name, (shape, role) in checkpoint_spec(self.key_dim, self.value_dim, self.pix_feat_dim).items():
            *path, leaf = name.split('.')
            node = self
            for part in path:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            if role in ('bn_mean', 'bn_var', 'bn_count'):
                node.register_buffer(leaf, init[name].clone())
            else:
                node.register_parameter(leaf, nn.Parameter(init[name].clone(), requires_grad=False))
        self._engine = None
        self._engine_key = None
        # 'native' = hand-written sm_100a conv stack (default); 'torch' = cuDNN/ATen fp32 (debug / comparison only)
        self.conv_backend = os.environ.get('DEVA_B200_CONV', 'native')
        self.return_full_logits = True  # segment() also returns the up-sampled logits like the reference

    @property
    def prefers_nhwc(self) -> bool:
        return self.conv_backend == 'native'

    # ------------------------------------------------------------------ weights
    def load_weights(self, src_dict: Dict[str, torch.Tensor]) -> None:
        self.load_state_dict(src_dict)
        self._engine = None

    def _apply(self, fn, *args, **kwargs):  # .cuda() / .to() invalidate the folded tables
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    @property
    def engine(self) -> Engine:
        probe = self.key_proj.key_proj.weight
        key = (probe.device, probe.data_ptr(), probe._version)
        if self._engine is None or self._engine_key != key:
            if probe.device.type != 'cuda':
                raise RuntimeError('deva_b200: the network runs on a CUDA device only (call .cuda()); '
                                   'there is no CPU fallback')
            with torch.no_grad():
                cls = NativeEngine if self.conv_backend == 'native' else Engine
                self._engine = cls({k: v.detach() for k, v in self.state_dict().items()})
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ reference API
    def aggregate(self, prob: torch.Tensor, dim: int) -> torch.Tensor:
        return Engine.aggregate(prob, dim)

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor) -> Tuple[Iterable[torch.Tensor], torch.Tensor]:
        return self.engine.encode_image(image.float())

    @torch.no_grad()
    def transform_key(self, feat: torch.Tensor, *, need_sk: bool = True, need_ek: bool = True):
        return self.engine.transform_key(feat, need_sk, need_ek)

    @torch.no_grad()
    def encode_mask(self, image: torch.Tensor, ms_features: Iterable[torch.Tensor], h: torch.Tensor,
                    masks: torch.Tensor, *, is_deep_update: bool = True, chunk_size: int = -1):
        return self.engine.encode_mask(image.float(), ms_features, h, masks, deep_update=is_deep_update,
                                       chunk_size=chunk_size)

    @torch.no_grad()
    def segment(self, multi_scale_features, memory_readout, sensory, last_mask, *, selector=None,
                need_aux: bool = False, chunk_size: int = -1, update_sensory: bool = True,
                independent_objects: bool = False):
        """Returns (sensory, logits [1,K+1,H,W], prob [1,K+1,H,W]) like network.py:94-173 (inference branch)."""
        if need_aux:
            raise NotImplementedError('need_aux is a training-time head; this engine is inference only')
        sensory, logits = self.engine.decode(multi_scale_features, memory_readout, sensory, last_mask,
                                             update_sensory=update_sensory, chunk_size=chunk_size)
        if self.conv_backend == 'native' and selector is None and not independent_objects:
            full_logits, prob = self.engine.probabilities(logits, want_logits=self.return_full_logits)
            return sensory, full_logits, prob
        logits = logits.float()
        prob = torch.sigmoid(logits)
        if selector is not None:
            prob = prob * selector
        if independent_objects:
            # per-object softmax against its own background (network.py:148-162); like the reference, the second
            # return value is the up-sampled aggregated LOGITS [K,2,H,W], not their softmax
            k, h, w = prob.shape[1:]
            each_logits = self.aggregate(prob.view(k, 1, h, w), dim=1)
            each_logits = F.interpolate(each_logits, scale_factor=4, mode='bilinear', align_corners=False)
            each = F.softmax(each_logits, dim=1)
            background = each[:, 0].min(dim=0)[0]
            prob = torch.cat([background.unsqueeze(0), each[:, 1]], dim=0).unsqueeze(0)
            return sensory, each_logits, prob
        logits = self.aggregate(prob, dim=1)
        logits = F.interpolate(logits, scale_factor=4, mode='bilinear', align_corners=False)
        return sensory, logits, F.softmax(logits, dim=1)
