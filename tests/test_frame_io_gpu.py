"""Frame ingest / egress kernels against the torch ops the reference's reader / driver use."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_ingest_rgb8_bit_exact():
    from deva.inference.frame_io import IMAGENET_MEAN, IMAGENET_STD, frame_from_rgb8
    g = torch.Generator().manual_seed(0)
    frame = torch.randint(0, 256, (270, 481, 3), generator=g, dtype=torch.uint8)
    out = frame_from_rgb8(frame.pin_memory())
    # the reference's reader runs these on the CPU (true divisions; CUDA torch would multiply by 1/255 instead)
    x = frame.permute(2, 0, 1).float().div(255)  # torchvision ToTensor
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    ref = (x - mean) / std  # torchvision Normalize
    assert out.shape == ref.shape and torch.equal(out.cpu(), ref)


@pytest.mark.parametrize('size,flip', [(None, False), (None, True), ((270, 480), False), ((1080, 1920), True), ((77, 131), False)])
def test_prob_to_ids_matches_driver_post_step(size, flip):
    from deva.inference.frame_io import prob_to_ids
    from deva.inference.object_info import ObjectInfo
    from deva.inference.object_manager import ObjectManager
    g = torch.Generator(device='cuda').manual_seed(1)
    k, h, w = 5, 120, 216
    prob = torch.softmax(4 * torch.randn(k + 1, h // 8, w // 8, device='cuda', generator=g), 0)
    prob = F.interpolate(prob.unsqueeze(0), size=(h, w), mode='bilinear', align_corners=False)[0].contiguous()
    om = ObjectManager()
    om.add_new_objects([ObjectInfo(i) for i in (7, 3, 200, 41, 9)])
    ref = prob
    if size is not None:
        ref = F.interpolate(ref.unsqueeze(1), size, mode='bilinear', align_corners=False)[:, 0]
    if flip:
        ref = torch.flip(ref, dims=[-1])
    top2 = torch.topk(ref, 2, dim=0)[0]
    ref_ids = om.tmp_to_obj_cls(torch.argmax(ref, dim=0))
    for dtype in (torch.long, torch.uint8):
        ids = prob_to_ids(prob, om, size=size, flip=flip, dtype=dtype)
        assert ids.dtype == dtype and ids.shape == ref_ids.shape
        same = ids.long() == ref_ids
        if size is None:
            assert bool(same.all())
        else:  # interpolation arithmetic may differ in the last bit: only exact ties of the top two may flip
            assert bool(same[(top2[0] - top2[1]) > 1e-6].all())
            assert float(same.float().mean()) > 0.9999
