"""Common evaluation flags and model loader (API of deva/inference/eval_args.py:7-71)."""
from argparse import ArgumentParser

import torch

from deva.model.network import DEVA

_FLAGS = [
    ('--model', dict(default='./saves/DEVA-propagation.pth')),
    ('--output', dict(default=None)),
    ('--save_all', dict(action='store_true', help='Save all frames')),
    ('--amp', dict(action='store_true')),
    ('--key_dim', dict(type=int, default=64)),
    ('--value_dim', dict(type=int, default=512)),
    ('--pix_feat_dim', dict(type=int, default=512)),
    ('--disable_long_term', dict(action='store_true')),
    ('--max_mid_term_frames', dict(type=int, default=10, help='T_max in XMem, decrease to save memory')),
    ('--min_mid_term_frames', dict(type=int, default=5, help='T_min in XMem, decrease to save memory')),
    ('--max_long_term_elements', dict(type=int, default=10000,
                                      help='LT_max in XMem, increase if objects disappear for a long time')),
    ('--num_prototypes', dict(type=int, default=128, help='P in XMem')),
    ('--top_k', dict(type=int, default=30)),
    ('--mem_every', dict(type=int, default=5, help='r in XMem. Increase to improve running speed.')),
    ('--chunk_size', dict(type=int, default=-1,
                          help='Number of objects to process in parallel as a batch; -1 for unlimited.')),
    ('--size', dict(type=int, default=480,
                    help='Resize the shorter side to this size. -1 to use original resolution.')),
]


def add_common_eval_args(parser: ArgumentParser):
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)


def get_model_and_config(parser: ArgumentParser):
    args = parser.parse_args()
    config = vars(args)
    config['enable_long_term'] = not config['disable_long_term']
    network = DEVA(config).cuda().eval()
    if args.model is not None:
        network.load_weights(torch.load(args.model, map_location='cuda'))
    else:
        print('No model loaded.')
    return network, config, args
