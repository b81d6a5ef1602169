"""GPU parity of the full DEVAInferenceCore.step path against reference-minted golden frames."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _core(cfg, sd, backend):
    from deva.inference.inference_core import DEVAInferenceCore
    from deva.model.network import DEVA
    net = DEVA(cfg)
    net.conv_backend = backend
    net = net.cuda().eval()
    net.load_weights({k: v.cuda() for k, v in sd.items()})
    return DEVAInferenceCore(net, cfg)


# 'native': every layer on the hand-written sm_100a kernels (fp16 activations, fp32 accumulate);
# 'torch': the same graphs through cuDNN fp32 (isolates the memory-read kernels).
@pytest.mark.parametrize('backend,tol', [('native', 2.5e-3), ('torch', 1e-3)])
def test_vos_clip_matches_reference(golden_dir, synthetic_sd, backend, tol):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, 'vos_steps.npz')).items()}
    meta = json.load(open(os.path.join(golden_dir, 'vos_steps.json')))
    np.random.seed(42)
    core = _core(meta['config'], synthetic_sd, backend)
    T = g['frames'].shape[0]
    worst = 0.0
    for t in range(T):
        img = g['frames'][t].cuda()
        if t == 0:
            p = core.step(img, g['mask0'].cuda(), [1, 2])
        elif t == 6:
            p = core.step(img, g['mask6'].cuda(), [7])
        else:
            p = core.step(img, end=(t == T - 1))
        ref = g[f'prob_{t:02d}']
        assert tuple(p.shape) == tuple(ref.shape)
        err = float((p.cpu() - ref).abs().max())
        worst = max(worst, err)
        sizes = {str(b): [core.memory.work_mem.size(b), core.memory.long_mem.size(b)]
                 for b in core.memory.work_mem.buckets}
        assert sizes == meta['sizes'][t], (t, sizes, meta['sizes'][t])
        # object-id indexing: confident pixels must agree exactly
        top2 = torch.topk(ref, 2, dim=0)[0]
        confident = (top2[0] - top2[1]) > 0.05
        assert bool((p.cpu().argmax(0)[confident] == ref.argmax(0)[confident]).all()), t
    print(f'[{backend}] max |prob - reference| over the clip:', worst)
    assert worst < tol, worst  # north_star target: 1e-3 max-abs vs the fp32 reference
    om = core.object_manager
    assert {t: o.id for t, o in om.tmp_id_to_obj.items()} == {1: 1, 2: 2, 3: 7}
    ids = om.tmp_to_obj_cls(torch.tensor([[0, 1], [2, 3]]).cuda())
    assert ids.cpu().tolist() == [[0, 1], [2, 7]]
