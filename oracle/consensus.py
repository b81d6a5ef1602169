"""CPU restatement of the reference's in-clip consensus (test infrastructure - never imported by the product).

    spatial_alignment                 deva/inference/consensus_associated.py:16-69
    established_association           deva/inference/consensus_associated.py:82-147
    auto_association                  deva/inference/consensus_automatic.py:82-272
    solve_brute_force                 the program of consensus_automatic.py:28-79, by enumeration

Pinned by tests/golden/consensus_*.{npz,json}, minted from the reference itself by tests/golden/make_golden.py (the
reference's `pulp` solver is absent from the image; the minting script plugs the same enumeration into the
reference's `solve_with_pulp` hook, which the reference itself documents as a replaceable fallback solver).
"""
from collections import defaultdict
from typing import Callable, List, Tuple

import numpy as np
import torch

from oracle import memory_math as mm
from oracle import network as net
from oracle.core import crop_pad, pad_to_multiple


class Segment:
    """Plain stand-in for ObjectInfo (object_info.py:7-62): id + the meta lists that `merge` concatenates."""
    def __init__(self, id, category_id=None, isthing=None, score=None):
        self.id, self.category_ids, self.scores, self.isthing = id, [category_id], [score], isthing

    def merge(self, other):
        self.category_ids.extend(other.category_ids)
        self.scores.extend(other.scores)


def spatial_alignment(sd, src_image, src_mask, tar_image, config) -> torch.Tensor:
    """src_image/tar_image [3,H,W] (padded), src_mask [K,H,W] -> target probabilities [1,K+1,H,W]."""
    k, h, w = src_mask.shape
    src_ms, src_feat = net.encode_image(sd, src_image.unsqueeze(0))
    src_key, src_shr, _ = net.transform_key(sd, src_feat)
    tar_ms, tar_feat = net.encode_image(sd, tar_image.unsqueeze(0))
    tar_key, _, tar_sel = net.transform_key(sd, tar_feat)
    sensory = torch.zeros(1, k, config['value_dim'], h // 16, w // 16)
    value, sensory = net.encode_mask(sd, src_image.unsqueeze(0), src_ms, sensory, src_mask.unsqueeze(0), deep_update=True)
    sim = mm.similarity(src_key[0].flatten(1), src_shr[0].flatten(), tar_key[0].flatten(1), tar_sel[0].flatten(1))
    aff = mm.dense_affinity(sim, config['top_k'])
    ro = mm.readout(aff, value[0].flatten(0, 1).flatten(1)).view(1, k, config['value_dim'], h // 16, w // 16)
    _, _, prob = net.segment(sd, tar_ms, ro, sensory, src_mask.unsqueeze(0), update_sensory=False)
    return prob


def established_association(sd, time_indices, images, masks, config, scores=None) -> Tuple[int, torch.Tensor]:
    padded = [pad_to_multiple(im, 16) for im in images]
    pads = padded[0][1]
    images = [p[0] for p in padded]
    masks = [pad_to_multiple(m, 16)[0] for m in masks]
    use_score = scores is not None
    scores = torch.softmax(torch.Tensor(scores if use_score else [1 for _ in time_indices]) * 2, dim=0).tolist()
    best, key = float('-inf'), None
    for i, (m, s) in enumerate(zip(masks, scores)):
        objective = s if use_score else float((m > 0.8).float().mean())
        if objective > best:
            best, key = objective, i
    key_score = scores[key] if use_score else scores[0]
    total = masks[key] * key_score
    for i, (im, m, s) in enumerate(zip(images, masks, scores)):
        if i != key:
            total = total + spatial_alignment(sd, im, m, images[key], config)[0, 1:] * s
    return time_indices[key], crop_pad(total, pads)


def solve_brute_force(pairwise_iou: np.ndarray, indicator: np.ndarray, total: int) -> List[bool]:
    """Enumerate all selections in ascending bitmask order; keep the first one that is better by more than 1e-9."""
    w = [float(pairwise_iou[:, i].sum() * 2) - 1.0 for i in range(total)]
    conflicts = [(i, j) for i in range(total) for j in range(i + 1, total) if indicator[i, j]]
    best_v, best_x = 0.0, 0
    for x in range(1, 1 << total):
        if any((x >> i) & 1 and (x >> j) & 1 for i, j in conflicts):
            continue
        v = sum(w[i] for i in range(total) if (x >> i) & 1)
        if v > best_v + 1e-9:
            best_v, best_x = v, x
    return [bool((best_x >> i) & 1) for i in range(total)]


def auto_association(frames: List[Tuple[int, torch.Tensor, torch.Tensor, List[Segment]]], keyframe_selection: str,
                     align: Callable, solver: Callable = solve_brute_force):
    """frames: (ti, image [3,H,W], id map [H,W] long, segments).  ``align(src_ti, src_image, src_mask, tar_ti,
    tar_image) -> [1,K+1,H,W]`` on padded tensors.  Returns (keyframe ti, id mask, [(id, category_ids, scores)])."""
    tis = [f[0] for f in frames]
    images, masks, pads = [], [], None
    for _, image, mask, _ in frames:
        image, pads = pad_to_multiple(image, 16)
        images.append(image)
        masks.append(pad_to_multiple(mask, 16)[0])
    next_id, info, per_frame, mappings = 0, {}, defaultdict(list), []
    for i, (_, _, _, segs) in enumerate(frames):
        one_hot, mapping = [], {}
        for si, seg in enumerate(segs):
            next_id += 1
            s = Segment(next_id)
            s.category_ids, s.scores, s.isthing = seg.category_ids, seg.scores, seg.isthing
            info[next_id] = s
            one_hot.append(masks[i] == seg.id)
            mapping[si] = next_id
            per_frame[i].append(s)
        masks[i] = torch.stack(one_hot, 0).float() if one_hot else None
        mappings.append(mapping)
    key_i = {'last': len(tis) - 1, 'first': 0, 'middle': (len(tis) + 1) // 2}[keyframe_selection]
    projected, area, pixels = [], {}, {}
    for ti, image, mask, mapping in zip(tis, images, masks, mappings):
        if mask is None:
            projected.append(None)
            continue
        if ti == tis[key_i]:
            prob = torch.cat([torch.ones_like(masks[key_i][0:1]) * 0.5, masks[key_i]], 0)
        else:
            prob = align(ti, image, mask, tis[key_i], images[key_i])[0]
        channel = torch.argmax(crop_pad(prob, pads), dim=0)
        remapped = torch.zeros_like(channel)
        for c, oid in mapping.items():
            hit = channel == (c + 1)
            remapped[hit] = oid
            area[oid], pixels[oid] = int(hit.sum()), hit
        projected.append(remapped)
    total = next_id
    if total == 0:
        return tis[key_i], torch.zeros_like(frames[0][2]), []
    table, iou = defaultdict(list), np.zeros((total, total), dtype=np.float32)
    for i in range(len(tis)):
        for j in range(i + 1, len(tis)):
            if projected[i] is None or projected[j] is None:
                continue
            for status in (None, False, True):
                taken = set()
                for a in per_frame[i]:
                    if a.isthing != status:
                        continue
                    for b in per_frame[j]:
                        if b.isthing != status or b.id in taken:
                            continue
                        inter = int(((projected[i] == a.id) & (projected[j] == b.id)).sum())
                        if inter == 0:
                            continue
                        v = inter / (area[a.id] + area[b.id] - inter)
                        if v > 0.5:
                            table[a.id].append(b.id)
                            table[b.id].append(a.id)
                            taken.add(b.id)
                            iou[a.id - 1, b.id - 1] = v
                            break
    iou = iou + iou.T
    indicator = iou > 0.49
    picked = solver(iou * indicator, indicator, total)
    out = torch.zeros_like(frames[0][2])
    chosen, infos = {}, []
    for c, sel in enumerate(picked):
        if sel:
            oid = c + 1
            chosen[oid] = area[oid]
            for other in table[oid]:
                info[oid].merge(info[other])
            infos.append(info[oid])
    for oid, _ in sorted(chosen.items(), key=lambda kv: kv[1], reverse=True):
        out[pixels[oid]] = oid
    return tis[key_i], out, [(s.id, list(s.category_ids), list(s.scores)) for s in infos]
