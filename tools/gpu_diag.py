"""GPU bring-up diagnostics: runs each check in its own subprocess (so one hung or crashing
kernel cannot hide the others) and appends JSON lines to gpurun_out/diag.jsonl.

    python tools/gpu_diag.py            # all groups
    python tools/gpu_diag.py conv_simt conv_fp16x3 ...   # selected groups
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'gpurun_out')

CONV_CASES = [
    # name, NB, C, H, W, Cout, k, stride, pad, res_mode, relu, force_im2col, block_n
    ('1x1_64_256_plain', 2, 64, 56, 56, 256, 1, 1, 0, 0, 1, 0, 0),
    ('1x1_256_64_tailM', 1, 256, 30, 30, 64, 1, 1, 0, 0, 1, 0, 0),
    ('1x1_res', 2, 64, 28, 28, 256, 1, 1, 0, 1, 1, 0, 0),
    ('1x1_up2x', 2, 512, 14, 14, 256, 1, 1, 0, 2, 0, 0, 0),
    ('1x1_as_im2col', 2, 128, 28, 28, 128, 1, 1, 0, 0, 0, 1, 0),
    ('1x1_s2', 2, 256, 56, 56, 512, 1, 2, 0, 0, 0, 0, 0),
    ('3x3_s1_64', 2, 64, 56, 56, 64, 3, 1, 1, 0, 1, 0, 0),
    ('3x3_s1_256', 1, 256, 14, 14, 256, 3, 1, 1, 0, 0, 0, 0),
    ('3x3_s2_128', 2, 128, 56, 56, 128, 3, 2, 1, 0, 1, 0, 0),
    ('3x3_7x7', 3, 512, 7, 7, 512, 3, 1, 1, 0, 1, 0, 0),
    ('1x1_bn64', 2, 256, 28, 28, 256, 1, 1, 0, 0, 0, 0, 64),
    ('1x1_bn128', 2, 256, 28, 28, 256, 1, 1, 0, 0, 0, 0, 128),
    ('1x1_bigK', 1, 2048, 7, 7, 512, 1, 1, 0, 0, 1, 0, 0),
]


def emit(rec):
    os.makedirs(OUT, exist_ok=True)
    rec['t'] = round(time.time(), 1)
    with open(os.path.join(OUT, 'diag.jsonl'), 'a') as f:
        f.write(json.dumps(rec) + '\n')
    print(json.dumps(rec), flush=True)


def group_conv(engine):
    import torch
    import torch.nn.functional as F
    from mcgaze_b200 import lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for (name, NB, C, H, W, Cout, k, stride, pad, res_mode, relu, fim, bn) in CONV_CASES:
        g = torch.Generator(device='cpu').manual_seed(hash(name) % 1000)
        x = torch.randn(NB, C, H, W, generator=g).cuda()
        w = (torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5).cuda()
        b = torch.randn(Cout, generator=g).cuda()
        P = (H + 2 * pad - k) // stride + 1
        Q = (W + 2 * pad - k) // stride + 1
        res = None
        if res_mode == 1:
            res = torch.randn(NB, Cout, P, Q, generator=g).cuda()
        elif res_mode == 2:
            res = torch.randn(NB, Cout, P // 2, Q // 2, generator=g).cuda()
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)
        if res_mode == 1:
            ref = ref + res.double()
        elif res_mode == 2:
            ref = ref + F.interpolate(res.double(), size=(P, Q), mode='nearest')
        if relu:
            ref = ref.relu()
        try:
            out = lib.debug_conv(engine, x, w, stride, pad, bias=b, res=res, res_mode=res_mode, relu=bool(relu),
                                 force_im2col=bool(fim), force_block_n=bn)
            torch.cuda.synchronize()
            err = (out.double() - ref).abs()
            emit({'group': 'conv_' + engine, 'case': name, 'max_abs_err': err.max().item(),
                  'mean_abs_err': err.mean().item(), 'ref_absmax': ref.abs().max().item(),
                  'nan': bool(torch.isnan(out).any().item())})
        except Exception as e:  # noqa
            emit({'group': 'conv_' + engine, 'case': name, 'error': repr(e)[:500]})


def group_forward(precision, B=1, T=7, hw=224, seed=0):
    import torch
    from mcgaze_b200 import lib
    from oracle import mcgaze_oracle as O
    sd = O.make_state_dict(0)
    img = O.make_clip(seed, B * T, hw, hw)
    taps = {}
    ref = O.forward(sd, img, clip_length=T, hk=O.Hooks(tap=lambda n, t: taps.__setitem__(n, t.clone())))
    eng = lib.Engine(sd, 0, precision)
    eng.set_option('keep_intermediates', 1)
    out = eng.forward(img.cuda(), clip_length=T)
    torch.cuda.synchronize()
    rec = {'group': 'forward_' + precision, 'B': B, 'T': T, 'hw': hw, 'launches': eng.last_launch_count}
    names = ['stem', 'pool', 'layer1.0', 'layer1.2', 'layer2.0', 'layer2.3', 'layer3.0', 'layer3.5', 'layer4.0',
             'layer4.2', 'fpn0', 'fpn1', 'fpn2', 'fpn3']
    for n in names:
        try:
            got = eng.intermediate(n).cpu()
            r = taps[n]
            rec['im_' + n] = [float((got - r).abs().max()), float(r.abs().max())]
        except Exception as e:  # noqa
            rec['im_' + n] = repr(e)[:200]
    for s in range(4):
        try:
            got = eng.intermediate(f'stage{s}.roi_feat').cpu()          # [R,49,256]
            r = taps[f'stage{s}.roi_feat'].flatten(2).permute(0, 2, 1)  # [R,256,7,7] -> [R,49,256]
            rec[f'im_stage{s}.roi_feat'] = [float((got - r).abs().max()), float(r.abs().max())]
            got = eng.intermediate(f'stage{s}.obj').cpu()
            r = taps[f'roi_head.bbox_head.{s}.obj']
            rec[f'im_stage{s}.obj'] = [float((got - r).abs().max()), float(r.abs().max())]
            got = eng.intermediate(f'stage{s}.attn').cpu()
            r = taps[f'roi_head.bbox_head.{s}.attn']
            rec[f'im_stage{s}.attn'] = [float((got - r).abs().max()), float(r.abs().max())]
            got = eng.intermediate(f'stage{s}.boxes').cpu()
            r = taps[f'stage{s}.boxes']
            rec[f'im_stage{s}.boxes'] = [float((got - r).abs().max()), float(r.abs().max())]
        except Exception as e:  # noqa
            rec[f'im_stage{s}'] = repr(e)[:200]
    g = out['gaze'].cpu()
    keys = ['gaze_score', 'face_gaze_score', 'eyes_gaze_score', 'head_gaze_score']
    for i, k in enumerate(keys):
        yp = O.vector_to_yaw_pitch(g[:, i])
        ypr = O.vector_to_yaw_pitch(ref[k])
        d = (yp - ypr).abs()
        d = torch.minimum(d, 2 * torch.pi - d)
        rec['yawpitch_err_' + k] = float(d.max())
        rec['vec_err_' + k] = float((g[:, i] - ref[k]).abs().max())
    rec['boxes_err'] = float((out['boxes'].cpu() - ref['boxes']).abs().max())
    rec['scores_err'] = float((out['scores'].cpu() - ref['scores']).abs().max())
    emit(rec)


def group_bench(precision, B=32, T=7, hw=224, steps=5, graph=0):
    import torch
    from mcgaze_b200 import lib
    from oracle import mcgaze_oracle as O
    sd = O.make_state_dict(0)
    eng = lib.Engine(sd, 0, precision)
    img = torch.randn(B * T, 3, hw, hw, device='cuda')
    out = eng.forward(img, clip_length=T)
    eng.set_graph_mode(bool(graph))
    for _ in range(3):
        eng.forward_into(img, T, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward_into(img, T, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    emit({'group': 'bench_' + precision, 'B': B, 'T': T, 'hw': hw, 'graph': graph, 'ms_per_step': ms,
          'clips_per_s': B / ms * 1e3, 'launches': eng.last_launch_count,
          'tflops_algorithmic': B * 99.55e9 / (ms * 1e-3) / 1e12})


GROUPS = {
    'conv_simt': lambda: group_conv('simt'),
    'conv_fp16x3': lambda: group_conv('fp16x3'),
    'conv_fp16': lambda: group_conv('fp16'),
    'forward_simt': lambda: group_forward('simt'),
    'forward_fp16x3': lambda: group_forward('fp16x3'),
    'forward_fp16': lambda: group_forward('fp16'),
    'forward_fp16x3_b2t3': lambda: group_forward('fp16x3', B=2, T=3, seed=3),
    'bench_simt_b2': lambda: group_bench('simt', B=2, steps=2),
    'bench_fp16x3': lambda: group_bench('fp16x3'),
    'bench_fp16': lambda: group_bench('fp16'),
    'bench_fp16x3_graph': lambda: group_bench('fp16x3', graph=1),
    'bench_fp16_graph': lambda: group_bench('fp16', graph=1),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == '--child':
        GROUPS[sys.argv[2]]()
        return
    groups = sys.argv[1:] or list(GROUPS)
    timeout = int(os.environ.get('DIAG_TIMEOUT', '300'))
    for gname in groups:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', gname], timeout=timeout,
                               capture_output=True, text=True)
            status = {'group': gname, 'rc': r.returncode, 'secs': round(time.time() - t0, 1)}
            if r.returncode != 0:
                status['stderr_tail'] = r.stderr[-1500:]
            emit(status)
            sys.stdout.write(r.stdout[-4000:])
        except subprocess.TimeoutExpired:
            emit({'group': gname, 'rc': 'timeout', 'secs': round(time.time() - t0, 1)})


if __name__ == '__main__':
    main()
