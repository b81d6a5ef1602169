"""Seeded scenario for the in-clip consensus fixtures, shared by make_golden.py (reference side), the oracle tests and
the GPU tests.  Pure torch - imports nothing from the reference, the product or the oracle."""
import torch

H, W = 90, 120  # padded to 96 x 128 by the consensus code (6 x 8 = 48 key positions >= top_k)
TIMES = [10, 11, 12, 13]

# Objects drift by (+2, -3) pixels per frame, which `shifted_alignment` undoes exactly; jitter in the boxes sets the IoUs.
# per frame: list of (segment id in that frame's id map, (y0, y1, x0, x1), category_id, isthing, score); later boxes
# overwrite earlier ones in the id map.
#   A: seen in frames 0, 1, 3 - mutually IoU > 0.5 (a triangle of conflicts: one representative survives)
#   B: frames 0, 1, 3 with IoU(0,1), IoU(1,3) > 0.5 > IoU(0,3): a chain - the middle detection has the most support
#   C, D: seen once (no support -> dropped);  S: "stuff" on top of A (never matched with a "thing")
#   N: isthing=None in frames 1 and 3 (matched among themselves; an exact tie -> lowest index)
DETECTIONS = [
    [(3, (10, 50, 40, 90), 1, True, 0.9), (5, (56, 86, 40, 80), 2, True, 0.8)],
    [(1, (13, 52, 37, 88), 1, True, 0.7), (2, (4, 30, 90, 118), 3, True, 0.6), (4, (58, 88, 43, 83), 2, True, 0.95),
     (6, (70, 88, 96, 116), None, None, None)],
    [],
    [(9, (18, 56, 32, 80), 4, False, 0.3), (7, (16, 58, 30, 81), 1, True, 0.5), (8, (62, 92, 45, 85), 2, True, 0.4),
     (2, (2, 14, 2, 30), 5, True, 0.2), (6, (74, 90, 90, 110), None, None, 0.1)],
]


def frames(seed: int = 21):
    """[(image [3,H,W] float32, id map [H,W] int64)] - temporally correlated noise, rectangles as detections.
    Later rectangles overwrite earlier ones where they overlap (an id map holds one id per pixel)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(3, H, W, generator=g)
    out = []
    for dets in DETECTIONS:
        image = base + 0.2 * torch.randn(3, H, W, generator=g)
        ids = torch.zeros(H, W, dtype=torch.long)
        for sid, (y0, y1, x0, x1), *_ in dets:
            ids[y0:y1, x0:x1] = sid
        out.append((image, ids))
    return out


def shifted_alignment(src_ti, src_image, src_mask, tar_ti, tar_image, *unused):
    """Stand-in for spatial_alignment with a known answer: every source mask moves by (2, -3) pixels per frame of
    distance to the target; the background channel is the constant 0.5 (like the keyframe's own projection)."""
    d = tar_ti - src_ti
    moved = torch.roll(src_mask, shifts=(2 * d, -3 * d), dims=(1, 2))
    return torch.cat([torch.ones_like(moved[0:1]) * 0.5, moved], dim=0).unsqueeze(0)
