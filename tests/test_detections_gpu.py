"""Semi-online session (incorporate_detection interleaved with step) against the reference-minted fixture.
Written after the round's GPU budget was spent: the oracle side is pinned on CPU (tests/test_oracle_golden.py), this
product-side comparison has NOT run on hardware yet - skipped unless DEVA_B200_TEST_EXPERIMENTAL=1."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('DEVA_B200_TEST_EXPERIMENTAL') != '1',
                                 reason='not yet validated on hardware; set DEVA_B200_TEST_EXPERIMENTAL=1')]


@pytest.mark.parametrize('backend,tol', [('native', 2.5e-3), ('torch', 1e-3)])
def test_detection_session_matches_reference(golden_dir, synthetic_sd, backend, tol):
    from deva.inference.inference_core import DEVAInferenceCore
    from deva.inference.object_info import ObjectInfo
    from deva.model.network import DEVA
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    spec = importlib.util.spec_from_file_location('consensus_scenario', os.path.join(golden_dir, 'consensus_scenario.py'))
    sc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sc)
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, 'detections.npz')).items()}
    meta = json.load(open(os.path.join(golden_dir, 'detections.json')))
    net = DEVA(meta['config'])
    net.conv_backend = backend
    net = net.cuda().eval()
    net.load_weights({k: v.cuda() for k, v in synthetic_sd.items()})
    np.random.seed(42)
    core = DEVAInferenceCore(net, meta['config'])
    frames = sc.detect_frames()
    for t, (frame, dets) in enumerate(zip(frames, sc.DETECT_SESSION)):
        if dets is None:
            p = core.step(frame.cuda(), end=(t == len(frames) - 1))
        else:
            ids = sc.merge_masks({d[0]: d[1] for d in dets}, sc.DETECT_HW).cuda()
            infos = [ObjectInfo(d[0], category_id=d[2], isthing=d[3], score=d[4]) for d in dets]
            p = core.incorporate_detection(frame.cuda(), ids, infos)
        want = meta['states'][t]
        objects = [[tt, o.id, o.poke_count, list(o.category_ids), list(o.scores)]
                   for tt, o in core.object_manager.tmp_id_to_obj.items()]
        assert objects == want['objects'], (t, objects, want['objects'])
        mem = core.memory
        sizes = {str(b): [mem.work_mem.size(b), mem.long_mem.size(b)] for b in mem.work_mem.buckets}
        assert sizes == want['sizes'], (t, sizes, want['sizes'])
        assert float((p.float().cpu() - g[f'prob_{t:02d}']).abs().max()) < tol, t
