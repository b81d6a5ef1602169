"""Micro-benchmark of deva_b200_conv2d on the decoder's dominant layer shapes (C3: 16 objects, 1080p)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tracking-anything-with-deva_b200'))
from deva import _native as nat  # noqa: E402
from deva.model import native_ops as ops  # noqa: E402

CASES = {
    # name: (batch, h, w, cin, cout, k, mode)
    'up84_c1': (16, 272, 480, 256, 256, 3, 'relu'),
    'up84_c1_actlo': (16, 272, 480, 256, 256, 3, 'actlo'),   # parity plan: activations as (hi, lo), two MMA passes
    'up84_c2_res': (16, 272, 480, 256, 256, 3, 'raw+res'),
    'up84_c2_head': (16, 272, 480, 256, 256, 3, 'raw+res+head'),
    'gru': (16, 68, 120, 512, 1536, 3, 'two'),
    'fuser_c2': (16, 68, 120, 512, 512, 3, 'raw+res'),
    'up168_c1': (16, 136, 240, 512, 256, 3, 'relu'),
    'ds_1x1': (16, 136, 240, 512, 256, 1, 'raw'),
    'res2_c3_precise': (1, 272, 480, 64, 256, 1, 'precise'),
    # shallow-K layers (epilogue / latency bound): sensory updater 1x1 convs, the mask stem as a 1x1 GEMM over 64 im2col columns
    'g16_1x1': (16, 68, 120, 512, 512, 1, 'raw'),
    'g8_1x1_res': (16, 68, 120, 256, 512, 1, 'raw+res'),
    'stem_1x1': (16, 544, 960, 64, 64, 1, 'relu'),
    # the sensory update as the product runs it: gate epilogue, 192-column tiles; and the plain epilogue on 192-column tiles
    'gru_gates': (16, 68, 120, 512, 1536, 3, 'gates'),
    'gru_n192': (16, 68, 120, 512, 1536, 3, 'two192'),
}


def run(name, iters=8):
    b, h, w, cin, cout, k, mode = CASES[name]
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(b, h, w, cin, device=dev, generator=g).half()
    two = mode in ('two', 'gates', 'two192')
    wgt = torch.randn(cout, cin * (2 if two else 1), k, k, device=dev, generator=g) / (cin * k * k)**0.5
    bias = torch.randn(cout, device=dev, generator=g)
    pc = ops.PackedConv(wgt, bias, 1, two_inputs=two, precise=(mode == 'precise'), act_lo=(mode == 'actlo'),
                        gates=(mode == 'gates'), nt_override=192 if mode == 'two192' else None)
    kw = {}
    if two:
        kw['x2'] = torch.randn_like(x)
    if mode == 'gates':
        kw['gate_h'] = kw['x2']
    if mode in ('precise', 'actlo'):
        kw['x_lo'] = (torch.randn_like(x) * 1e-3)
    if 'res' in mode:
        kw['res'] = torch.randn(b, h, w, cout, device=dev, generator=g).half()
    if 'head' in mode:
        kw['head_w'] = torch.randn(9, cout, device=dev, generator=g)
    want = dict(want_relu=True) if mode in ('relu', ) else (dict() if mode == 'gates' else dict(want_raw=True))
    if mode in ('precise', 'actlo'):
        want = dict(want_relu=True, want_lo=True)
    passes = 3 if mode == 'precise' else (2 if (two or mode == 'actlo') else 1)
    flops = 2.0 * b * h * w * cout * k * k * pc.cin_pad * passes
    for _ in range(3):
        ops.conv_ex(x, pc, **kw, **want)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_ex(x, pc, **kw, **want); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    return dict(case=name, ms=ms, tflops=flops / ms / 1e9, gflop=flops / 1e9)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default=','.join(CASES))
    a = ap.parse_args()
    nat.require_device()
    for c in a.cases.split(','):
        print(json.dumps(run(c)))
