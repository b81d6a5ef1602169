"""Checkpoint layout of the DEVA propagation network, as a flat table.

The drop-in boundary for weights is the flat ``state_dict`` the reference saves with
``torch.save`` and loads in deva/inference/eval_args.py:66-68 (keys ``pixel_encoder.*``,
``mask_encoder.*``, ``key_proj.*``, ``mask_decoder.*``).  This module rebuilds that table
(name -> shape, role) from the architecture description instead of from an ``nn.Module``
tree, so the engine can validate / allocate / synthesise checkpoints without the reference.

Roles: 'conv' (OIHW weight), 'linear' ([out,in] weight), 'bias', 'bn_gamma', 'bn_beta',
'bn_mean', 'bn_var', 'bn_count'.
"""
from collections import OrderedDict
from typing import Dict, Tuple

import torch

Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str]]"


class _Table:
    def __init__(self):
        self.rows: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def conv(self, name, cin, cout, k, bias):
        self.rows[name + '.weight'] = ((cout, cin, k, k), 'conv')
        if bias:
            self.rows[name + '.bias'] = ((cout, ), 'bias')

    def linear(self, name, cin, cout):
        self.rows[name + '.weight'] = ((cout, cin), 'linear')
        self.rows[name + '.bias'] = ((cout, ), 'bias')

    def bn(self, name, c):
        self.rows[name + '.weight'] = ((c, ), 'bn_gamma')
        self.rows[name + '.bias'] = ((c, ), 'bn_beta')
        self.rows[name + '.running_mean'] = ((c, ), 'bn_mean')
        self.rows[name + '.running_var'] = ((c, ), 'bn_var')
        self.rows[name + '.num_batches_tracked'] = ((), 'bn_count')


def _bottleneck_stage(t: _Table, prefix, cin, planes, blocks, stride):
    """ResNet-50 stage (reference resnet.py:78-152)."""
    for i in range(blocks):
        p = f'{prefix}.{i}'
        inp = cin if i == 0 else planes * 4
        t.conv(p + '.conv1', inp, planes, 1, False); t.bn(p + '.bn1', planes)
        t.conv(p + '.conv2', planes, planes, 3, False); t.bn(p + '.bn2', planes)
        t.conv(p + '.conv3', planes, planes * 4, 1, False); t.bn(p + '.bn3', planes * 4)
        if i == 0 and (stride != 1 or inp != planes * 4):
            t.conv(p + '.downsample.0', inp, planes * 4, 1, False); t.bn(p + '.downsample.1', planes * 4)


def _basic_stage(t: _Table, prefix, cin, planes, blocks, stride):
    """ResNet-18 stage (reference resnet.py:46-75,131-145)."""
    for i in range(blocks):
        p = f'{prefix}.{i}'
        inp = cin if i == 0 else planes
        t.conv(p + '.conv1', inp, planes, 3, False); t.bn(p + '.bn1', planes)
        t.conv(p + '.conv2', planes, planes, 3, False); t.bn(p + '.bn2', planes)
        if i == 0 and (stride != 1 or inp != planes):
            t.conv(p + '.downsample.0', inp, planes, 1, False); t.bn(p + '.downsample.1', planes)


def _fusion(t: _Table, prefix, cin, mid, cout):
    """GroupFeatureFusionBlock (reference group_modules.py:133-152, cbam.py:21-77)."""
    t.conv(prefix + '.block1.downsample', cin, mid, 1, True)
    t.conv(prefix + '.block1.conv1', cin, mid, 3, True)
    t.conv(prefix + '.block1.conv2', mid, mid, 3, True)
    t.linear(prefix + '.attention.ChannelGate.mlp.1', mid, mid // 16)
    t.linear(prefix + '.attention.ChannelGate.mlp.3', mid // 16, mid)
    t.conv(prefix + '.attention.SpatialGate.spatial.conv', 2, 1, 7, True)
    t.conv(prefix + '.block2.conv1', mid, cout, 3, True)
    t.conv(prefix + '.block2.conv2', cout, cout, 3, True)


def checkpoint_spec(key_dim: int = 64, value_dim: int = 512, pix_feat_dim: int = 512):
    """Ordered table of every tensor in a DEVA propagation checkpoint."""
    t = _Table()
    # pixel_encoder: ResNet-50 to layer3 + two 1x1 projections (big_modules.py:23-51)
    p = 'pixel_encoder'
    t.conv(p + '.conv1', 3, 64, 7, False); t.bn(p + '.bn1', 64)
    _bottleneck_stage(t, p + '.res2', 64, 64, 3, 1)
    _bottleneck_stage(t, p + '.layer2', 256, 128, 4, 2)
    _bottleneck_stage(t, p + '.layer3', 512, 256, 6, 2)
    t.conv(p + '.proj1', 1024, pix_feat_dim, 1, True)
    t.conv(p + '.proj2', 1024, pix_feat_dim, 1, True)
    # mask_encoder: ResNet-18 to layer3 on image+mask, fuser, deep sensory GRU (:54-127)
    p = 'mask_encoder'
    t.conv(p + '.conv1', 4, 64, 7, False); t.bn(p + '.bn1', 64)
    _basic_stage(t, p + '.layer1', 64, 64, 2, 1)
    _basic_stage(t, p + '.layer2', 64, 128, 2, 2)
    _basic_stage(t, p + '.layer3', 128, 256, 2, 2)
    _fusion(t, p + '.fuser', pix_feat_dim + 256, value_dim, value_dim)
    t.conv(p + '.sensory_update.transform', value_dim * 2, value_dim * 3, 3, True)
    # key_proj (modules.py:60-78)
    t.conv('key_proj.key_proj', pix_feat_dim, key_dim, 3, True)
    t.conv('key_proj.d_proj', pix_feat_dim, 1, 3, True)
    t.conv('key_proj.e_proj', pix_feat_dim, key_dim, 3, True)
    # mask_decoder (big_modules.py:130-145)
    p = 'mask_decoder'
    _fusion(t, p + '.fuser', 512 + value_dim, value_dim, value_dim)
    t.conv(p + '.sensory_compress', value_dim + 1, value_dim, 1, True)
    t.conv(p + '.sensory_update.g16_conv', value_dim, 512, 1, True)
    t.conv(p + '.sensory_update.g8_conv', 256, 512, 1, True)
    t.conv(p + '.sensory_update.g4_conv', 257, 512, 1, True)
    t.conv(p + '.sensory_update.transform', 512 + 512, 512 * 3, 3, True)
    t.conv(p + '.decoder_feat_proc.transforms.0', 512, value_dim, 1, True)
    t.conv(p + '.decoder_feat_proc.transforms.1', 256, 256, 1, True)
    t.conv(p + '.up_16_8.out_conv.downsample', value_dim, 256, 1, True)
    t.conv(p + '.up_16_8.out_conv.conv1', value_dim, 256, 3, True)
    t.conv(p + '.up_16_8.out_conv.conv2', 256, 256, 3, True)
    t.conv(p + '.up_8_4.out_conv.conv1', 256, 256, 3, True)
    t.conv(p + '.up_8_4.out_conv.conv2', 256, 256, 3, True)
    t.conv(p + '.pred', 256, 1, 3, True)
    t.conv(p + '.sensory_linear_pred.projection', value_dim, 512 + 1, 1, True)  # training-only head
    return t.rows


def synthetic_state_dict(seed: int = 0, key_dim: int = 64, value_dim: int = 512,
                         pix_feat_dim: int = 512) -> Dict[str, torch.Tensor]:
    """Deterministic random checkpoint with the real architecture (no network download).

    Scaled so activations stay O(1) through ~50 layers (safe in fp16) and the memory read is
    well conditioned: conv/linear weights ~ N(0, sqrt(1/fan_in)), BN statistics near identity
    with mild random affine terms, the last BN of every residual branch damped so residual sums
    do not grow per block, keys ~ unit variance, shrinkage in [1, ~3], logits with a few units
    of spread.  Drawn from a CPU generator: the same seed gives the same tensors everywhere.
    """
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    out = OrderedDict()
    for name, (shape, role) in checkpoint_spec(key_dim, value_dim, pix_feat_dim).items():
        if role in ('conv', 'linear'):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            std = (1.0 / fan_in)**0.5
            if name.endswith('.conv1.weight') and name.count('.') == 2:
                std *= 1.4  # stems see zero-mean images
            if 'd_proj' in name:
                std *= 0.5
            if name.endswith('pred.weight'):
                std *= 1.5
            if 'sensory_update.transform' in name:
                std *= 1.5
            out[name] = torch.randn(shape, generator=g) * std
        elif role == 'bias':
            out[name] = torch.randn(shape, generator=g) * 0.05
        elif role == 'bn_gamma':
            last = name.endswith('bn3.weight') or (name.endswith('bn2.weight') and 'mask_encoder' in name)
            out[name] = (0.3 if last else 1.0) * (1.0 + 0.1 * torch.randn(shape, generator=g))
        elif role == 'bn_beta':
            out[name] = 0.05 * torch.randn(shape, generator=g)
        elif role == 'bn_mean':
            out[name] = 0.05 * torch.randn(shape, generator=g)
        elif role == 'bn_var':
            out[name] = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif role == 'bn_count':
            out[name] = torch.zeros((), dtype=torch.int64)
    return out
