#!/usr/bin/env python
"""Benchmark of the DEVA propagation hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2]

Workload (BASELINE.json configs[2], "c3"): synthetic 1080p frames (padded 1088x1920, Q = 8160 query
positions), 16 objects, memory bank pre-filled to 10 000 slots, full encode -> read -> decode per frame
through ``DEVAInferenceCore.step``; every 5th frame is a memory frame (value encoder + bank append,
``mem_every=5``) after which the bank is clamped back to 10 000 slots so the configuration stays the named
one.  One step = one frame.  Metric = propagation FPS (whole job, all ranks).  N > 1: one process per GPU,
each rank propagates its own clip (clip-parallel, weak scaling, NCCL barrier only - BASELINE configs[3]).

Printed JSON line: metric/value/unit/... plus
  roofline      fused affinity path (pack_query + similarity/top-k/softmax + readout GEMM), algorithmic
                FLOPs 2*N*Q*2CK + 2*K*CV*N*Q per frame / CUDA-event time per frame, vs measured bf16 peak;
  e2e           same FPS through the public API with host frames (pinned H2D of every frame, D2H of the id mask);
  cpu_baseline  the CPU oracle (port of the reference, oracle/) on this box's host cores, bounded sample;
  clocks        SM clock / throttle reasons sampled with nvidia-smi during the timed region.
``--impl reference`` times the CPU oracle instead (the reference is pure Python/PyTorch; its CPU path is what
``oracle/`` restates and pins to reference-minted fixtures).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'tracking-anything-with-deva_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CK, CV, TOP_K = 64, 512, 30
WORKLOADS = {
    'c3': dict(name='c3: 1080p, 16 objects, 10k memory slots, full encode->read->decode', h=1080, w=1920, k=16,
               n=10000),
    'c2': dict(name='c2: 480p, 5 objects, 2k memory slots', h=480, w=854, k=5, n=2000),
}
METRIC = 'propagation FPS @1080p, 10k-mem, 16 obj; affinity GEMM TFLOPS vs bf16 peak'


def base_config():
    return dict(key_dim=CK, value_dim=CV, pix_feat_dim=512, mem_every=5, enable_long_term=True, chunk_size=-1,
                top_k=TOP_K, enable_long_term_count_usage=True, max_mid_term_frames=10, min_mid_term_frames=5,
                num_prototypes=128, max_long_term_elements=10000)


def synth_frames(wl, count, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(3, wl['h'], wl['w'], generator=g)
    return torch.stack([base + 0.2 * torch.randn(3, wl['h'], wl['w'], generator=g) for _ in range(count)])


def synth_mask(wl):
    """K rectangles on a grid, ids 1..K."""
    k, h, w = wl['k'], wl['h'], wl['w']
    cols = 4 if k > 4 else k
    rows = (k + cols - 1) // cols
    m = torch.zeros(h, w, dtype=torch.long)
    for i in range(k):
        r, c = divmod(i, cols)
        y0, x0 = int((r + 0.15) * h / rows), int((c + 0.15) * w / cols)
        m[y0:y0 + int(0.6 * h / rows), x0:x0 + int(0.6 * w / cols)] = i + 1
    return m


def read_flops(wl, q):
    return 2.0 * wl['n'] * q * 2 * CK + 2.0 * wl['k'] * CV * wl['n'] * q


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                     f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ ours
class Clip:
    """One clip on one GPU: network, core, bank pre-filled to wl['n'] slots."""
    def __init__(self, wl, device, seed):
        from deva import _native as nat
        from deva.inference.inference_core import DEVAInferenceCore
        from deva.model.network import DEVA
        from deva.model.param_spec import synthetic_state_dict
        nat.require_device()
        self.nat, self.wl, self.device = nat, wl, device
        cfg = base_config()
        net = DEVA(cfg).to(device).eval()
        net.load_weights({k: v.to(device) for k, v in synthetic_state_dict(seed=0).items()})
        self.core = DEVAInferenceCore(net, cfg)
        self.frames_host = synth_frames(wl, 5, seed).pin_memory()
        self.frames_dev = self.frames_host.to(device)
        ids = list(range(1, wl['k'] + 1))
        self.core.step(self.frames_dev[0], synth_mask(wl).to(device), ids)  # first frame -> Q memory tokens
        mem = self.core.memory
        bank = next(iter(mem._banks.values()))
        self.q = mem.HW
        extra = wl['n'] - bank.work_size
        assert extra >= 0, 'bank already larger than the configured slot count'
        if extra > 0:  # random-init top-up to exactly n slots (BASELINE.md section 4 generator)
            g = torch.Generator(device=device).manual_seed(seed + 1)
            key = torch.randn(1, CK, extra, 1, device=device, generator=g)
            shr = 1 + torch.rand(1, 1, extra, 1, device=device, generator=g)
            sel = torch.sigmoid(torch.randn(1, CK, extra, 1, device=device, generator=g))
            val = torch.randn(1, wl['k'], CV, extra, 1, device=device, generator=g)
            mem.add_memory(key, shr, val, ids, selection=sel)
        self.bank, self.mark = bank, bank.hi
        assert bank.work_size == wl['n']
        self.i = 0

    def clamp(self):
        if self.core.last_mem_ti == self.core.curr_ti:  # a memory frame was just appended
            self.bank.hi = self.mark

    def step_resident(self):
        p = self.core.step(self.frames_dev[self.i % 5])
        self.clamp()
        self.i += 1
        return p

    def step_e2e(self):
        """What a user of the public API does per frame (evaluation/eval_vos.py:138-198)."""
        img = self.frames_host[self.i % 5].to(self.device, non_blocking=True)
        p = self.core.step(img)
        ids = self.core.object_manager.tmp_to_obj_cls(torch.argmax(p, dim=0)).to(torch.uint8)
        host = ids.cpu()
        self.clamp()
        self.i += 1
        return host


    def step_e2e_fused_io(self):
        """Same, with the decoded uint8 frame uploaded as is and the ingest / egress kernels of SURVEY 8(f)-3
        (deva.inference.frame_io): normalise on the device, fused argmax + id remap -> uint8 id map."""
        from deva.inference.frame_io import frame_from_rgb8, prob_to_ids
        if not hasattr(self, 'frames_u8'):
            mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
            u8 = ((self.frames_host * std + mean) * 255).round().clamp(0, 255).to(torch.uint8)
            self.frames_u8 = u8.permute(0, 2, 3, 1).contiguous().pin_memory()
        img = frame_from_rgb8(self.frames_u8[self.i % 5], device=self.device)
        p = self.core.step(img)
        host = prob_to_ids(p, self.core.object_manager, dtype=torch.uint8).cpu()
        self.clamp()
        self.i += 1
        return host


def timed(fn, steps, dist_on):
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), 0.0)
    if dist_on:
        t = torch.tensor([ms, wall], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms, wall = float(t[0]), float(t[1])
    return ms, wall


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    dist_on = world > 1
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if dist_on:
        dist.init_process_group('nccl', device_id=device)
    torch.backends.cudnn.benchmark = True
    wl = WORKLOADS[args.workload]
    clip = Clip(wl, device, seed=100 + rank)
    nat = clip.nat
    for _ in range(max(args.warmup, 3)):
        clip.step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    clip.core.memory.read_events = []
    l0 = nat.launch_count()
    ms, wall = timed(clip.step_resident, args.steps, dist_on)
    launches = nat.launch_count() - l0
    ev = clip.core.memory.read_events
    clip.core.memory.read_events = None
    read_ms = sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        clip.step_e2e()
    ms_e2e, _ = timed(clip.step_e2e, args.steps, dist_on)
    for _ in range(2):
        clip.step_e2e_fused_io()
    ms_e2e_io, _ = timed(clip.step_e2e_fused_io, args.steps, dist_on)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)' if peaks else \
            'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        flops = read_flops(wl, clip.q)
        achieved = flops / (read_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get(args.workload)  # readout_sparse_kernel
        except Exception:
            pass
        h2d = int(clip.frames_host[0].numel() * 4)
        d2h = int(wl['h'] * wl['w'])
        out = {
            'metric': METRIC, 'value': world * args.steps / (ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16 operands / f32 accumulate '
            '(tcgen05 memory read + conv stack); key path split-f16x3 (~f32)', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'frame': [wl['h'], wl['w']], 'query_positions': clip.q,
                       'objects': wl['k'], 'memory_slots': wl['n'], 'mem_every': 5, 'top_k': TOP_K,
                       'parallelism': f'clip-parallel x{world}' if world > 1 else 'single clip',
                       'l2': 'per-step working set (activations > 2 GB) exceeds the 126 MB L2; no explicit flush',
                       'weights': 'synthetic_state_dict(seed=0), real architecture (69.2 M parameters)'},
            'roofline': {'kernel': 'fused affinity path: pack_query + sim_topk(tcgen05 fp16x3) + merge + bucket + readout_sparse(tcgen05, '
                                   'affinity tiles built in smem); traffic = DRAM bytes of readout_sparse_kernel (ncu)',
                         'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'ms_per_launch': read_ms, 'flops_per_launch': flops},
            'e2e': {'value': world * args.steps / (ms_e2e * 1e-3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            'e2e_fused_io': {'value': world * args.steps / (ms_e2e_io * 1e-3), 'unit': 'frames/s',
                             'h2d_bytes_per_step': int(wl['h'] * wl['w'] * 3), 'd2h_bytes_per_step': d2h,
                             'what': 'uint8 frame upload + on-device normalise; fused argmax + id remap (deva.inference.frame_io)'},
            'gpu_launches': int(launches), 'wall_ms_per_step': wall / args.steps, 'clocks': clocks,
            'cpu_baseline': cpu_baseline(wl, budget_s=25.0),
        }
        if world == 1 and not args.no_torch_baseline:
            del clip  # give the stock-PyTorch pass the whole device
            torch.cuda.empty_cache()
            try:
                out['torch_gpu_baseline'] = torch_gpu_baseline(wl, device)
            except Exception as exc:  # reported extra, never fatal for the bench line
                out['torch_gpu_baseline'] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
        print(json.dumps(out))
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- CPU oracle
def _oracle_stages(wl, budget_s, device, k_s):
    """One frame of the reference algorithm (the oracle's fp32 PyTorch restatement) on ``device``, stage by stage.
    Returns wall-clock seconds per stage; per-object stages run on ``k_s`` objects."""
    from deva.model.param_spec import synthetic_state_dict
    from oracle import memory_math as mm
    from oracle import network as onet
    from oracle.core import pad_to_multiple
    cuda = torch.device(device).type == 'cuda'

    def clock():
        if cuda:
            torch.cuda.synchronize()
        return time.perf_counter()

    sd = {name: v.to(device) for name, v in synthetic_state_dict(seed=0).items()}
    t_all = time.perf_counter()
    with torch.no_grad():
        img, _ = pad_to_multiple(synth_frames(wl, 1, 7)[0], 16)
        img = img.unsqueeze(0).to(device)
        t0 = clock()
        ms, feat = onet.encode_image(sd, img)
        key, shr, sel = onet.transform_key(sd, feat)
        t_shared = clock() - t0
        h, w = key.shape[-2:]
        n = wl['n']
        g = torch.Generator().manual_seed(0)
        mk, msh = torch.randn(CK, n, generator=g).to(device), (1 + torch.rand(n, generator=g)).to(device)
        mv = torch.randn(k_s * CV, n, generator=g).to(device)
        t0 = clock()
        sim = mm.similarity(mk, msh, key[0].flatten(1), sel[0].flatten(1))
        aff = mm.dense_affinity(sim, TOP_K)
        t_aff = clock() - t0
        t0 = clock()
        ro = mm.readout(aff, mv)
        t_ro = clock() - t0
        del sim, aff
        masks = synth_mask(wl)
        masks = torch.stack([(masks == (i % wl['k']) + 1).float() for i in range(k_s)])
        masks, _ = pad_to_multiple(masks, 16)
        masks = masks.unsqueeze(0).to(device)
        sens = torch.zeros(1, k_s, CV, h, w, device=device)
        t0 = clock()
        onet.segment(sd, ms, ro.view(1, k_s, CV, h, w), sens, masks)
        t_seg = clock() - t0
        t_enc = None
        if time.perf_counter() - t_all < budget_s:
            t0 = clock()
            onet.encode_mask(sd, img, ms, sens, masks)
            t_enc = clock() - t0
    return {'encode': t_shared, 'affinity': t_aff, 'readout': t_ro, 'decode': t_seg, 'encode_mask': t_enc}


def _blend(st, scale):
    per_frame = st['encode'] + st['affinity'] + (st['readout'] + st['decode']) * scale
    if st['encode_mask'] is not None:
        per_frame += st['encode_mask'] * scale / 5.0
    return per_frame


def cpu_baseline(wl, budget_s):
    """Reference-algorithm frame rate on the host cores: the oracle's stages on a bounded sample.

    Object-independent stages run in full; per-object stages run on ``k_s`` of the K objects and are scaled by
    K/k_s (every per-object op is independent across objects)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    k, k_s = wl['k'], 1
    st = _oracle_stages(wl, budget_s, 'cpu', k_s)
    scale = k / k_s
    return {'value': 1.0 / _blend(st, scale), 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': f'1 frame {wl["h"]}x{wl["w"]}, N={wl["n"]}: encode_image+transform_key and similarity/top-k in '
                      f'full; readout, decoder' + (', value encoder (1 frame in 5)' if st['encode_mask'] is not None else '') +
                      f' on {k_s} of {k} objects, scaled x{scale:g}; oracle = fp32 PyTorch-CPU port of the reference',
            'stage_s': {'encode': st['encode'], 'affinity': st['affinity'], 'readout_per_obj': st['readout'],
                        'decode_per_obj': st['decode'], 'encode_mask_per_obj': st['encode_mask']}}


def torch_gpu_baseline(wl, device):
    """SURVEY 8(d) "GPU-side comparison": the same reference algorithm as stock PyTorch ops (cuDNN / cuBLAS fp32,
    PyTorch's default TF32 policy) on the same B200, all K objects in one batch; per-stage best of three warm passes.  A reported baseline only - none of it is on the product path."""
    k = wl['k']
    st = None
    for i in range(4):  # pass 0 pays cuDNN's algorithm search; keep the per-stage best of the others
        cur = _oracle_stages(wl, 1e9, device, k)
        torch.cuda.empty_cache()
        if i == 1:
            st = cur
        elif i > 1:
            st = {name: min(st[name], cur[name]) for name in st}
    return {'value': 1.0 / _blend(st, 1.0), 'unit': 'frames/s', 'kind': 'oracle ops on cuda (stock PyTorch fp32)',
            'tf32': {'cudnn': bool(torch.backends.cudnn.allow_tf32), 'matmul': bool(torch.backends.cuda.matmul.allow_tf32)},
            'sample': f'1 frame {wl["h"]}x{wl["w"]}, N={wl["n"]}, all {k} objects; value encoder weighted 1 frame in 5',
            'stage_s': st}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_baseline(wl, budget_s=40.0))
    best = max(vals, key=lambda v: v['value'])
    out = {'impl': 'reference', 'metric': METRIC, 'value': best['value'], 'unit': 'frames/s', 'n_gpus': world,
           'steps': len(vals), 'warmup': 0, 'ms_per_step': 1e3 / best['value'], 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': wl['name'], 'frame': [wl['h'], wl['w']], 'objects': wl['k'],
                      'memory_slots': wl['n']},
           'cpu_baseline': best,
           'e2e': {'value': best['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=list(WORKLOADS))
    ap.add_argument('--no-torch-baseline', action='store_true', help='skip the stock-PyTorch-on-GPU comparison pass')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
