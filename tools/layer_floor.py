"""CPU-only: per-launch roofline floors of the trunk + FPN against a committed ncu launch list.

    python tools/layer_floor.py [profiles/r02_launches_fp16c8_f.csv]          # round-2 schedule (default)
    python tools/layer_floor.py profiles/r01_launches_fp16c8_g.csv r01        # round-1 schedule (52 separate convolutions)

For every tcgen05 launch of the trunk and the FPN of ONE bench step (ResNet-50 + FPN at 224^2, 32 clips x 7 frames, fp16c8)
prints the measured duration, the algorithmic FLOPs and bytes, and the floor

    max(bytes / HBM peak, 2 x FLOPs / tensor peak)

- fp16c8 spends two fp16-MMA units per algorithmic MMA; activations are 4 B / element (fp16 hi + e4m3
lo8 + e4m3 hi8), 3 B where a tensor is only read as a residual, FPN outputs 3 B (no hi8: RoIAlign reads them).  Round-2 schedule
(DESIGN.md section 1): conv3 + downsample of a layer's first block are ONE K-concatenated launch (the downsample tensor does
not exist), conv2 -> conv3 + identity of layer1.1-2 / layer2.1-3 are ONE launch (t2 does not exist).  Peaks from
MEASURED_PEAKS.json (fallback: the B200_PROFILING.md numbers).  The launch list is cold-cache and serialised (ncu): the per-
launch ratios are what to read, not the sum."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_launches_fp16c8_f.csv')
mode = sys.argv[2] if len(sys.argv) > 2 else ('r01' if '/r01_' in path.replace('\\', '/') else 'r02')
try:
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    HBM, TF = pk['hbm_gbs'] * 1e9, pk['bf16_tflops_sustained'] * 1e12
except Exception:
    HBM, TF = 6650e9, 1400e12
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
dur = [float(r['Metric Value'].replace(',', '')) / 1e3 for r in rows]          # us
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void mcg::', '').replace('mcg::', '') for r in rows]
# the list may hold several forwards (warm-up + profiled): keep the last one
starts = [i for i, n in enumerate(names) if n.startswith('stem_fused')]
assert starts, 'no stem_fused_kernel launch: not a launch list of the forward'
dur, names = dur[starts[-1]:], names[starts[-1]:]

NB = 224
# (label, expected kernel prefix, FLOPs, bytes)
launches = []


def add(label, kernel, M, N, K, in_bytes, out_bytes):
    launches.append((label, kernel, 2.0 * M * N * K, in_bytes + out_bytes + N * K * 4))


h, cin = 56, 64
for l, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
    for b in range(blocks):
        stride = 2 if (b == 0 and l > 0) else 1
        ho = h // stride
        pin, pout = NB * h * h, NB * ho * ho
        name = f'layer{l + 1}.{b}'
        add(f'{name}.conv1', 'umma', pin, planes, cin, pin * cin * 4, pin * planes * 4)
        fused_tail = mode == 'r02' and b > 0 and l < 2
        if fused_tail:
            # t1 (4 B) + identity (3 B) read, y (4 B) written; FLOPs of conv2 + conv3
            fl = 2.0 * pout * planes * 9 * planes + 2.0 * pout * 4 * planes * planes
            by = pin * planes * 4 + pout * 4 * planes * 3 + pout * 4 * planes * 4 + (9 * planes * planes + 4 * planes * planes) * 4
            launches.append((f'{name}.conv2+conv3+identity (fused tail)', 'bneck', fl, by))
        else:
            add(f'{name}.conv2', 'umma', pout, planes, 9 * planes, pin * planes * 4, pout * planes * 4)
            if b == 0 and mode == 'r02':
                # K-concatenated: t2 (4 B) + the block input read with the block's stride (4 B per pixel read), y written
                add(f'{name}.conv3+downsample (K-concatenated)', 'umma', pout, 4 * planes, planes + cin,
                    pout * planes * 4 + pout * cin * 4, pout * 4 * planes * 4)
            else:
                if b == 0:
                    add(f'{name}.downsample', 'umma', pout, 4 * planes, cin, pout * cin * 4, pout * 4 * planes * 3)
                add(f'{name}.conv3+identity', 'umma', pout, 4 * planes, planes, pout * planes * 4 + pout * 4 * planes * 3,
                    pout * 4 * planes * 4)
        cin, h = 4 * planes, ho
if mode == 'r02':
    for i, c in ((3, 2048), (2, 1024), (1, 512), (0, 256)):
        s = 56 >> i
        p = NB * s * s
        add(f'fpn lateral{i} (+ top-down add)' if i < 3 else f'fpn lateral{i}', 'umma', p, 256, c,
            p * c * 4 + (p // 4 * 256 * 3 if i < 3 else 0), p * 256 * 4)
    for i in range(4):
        s = 56 >> i
        p = NB * s * s
        add(f'fpn output{i} 3x3', 'umma', p, 256, 9 * 256, p * 256 * 4, p * 256 * 3)

print(f'{"launch":46s} {"kernel":>8s} {"us":>7s} {"GFLOP":>7s} {"TFLOP/s":>8s} {"MB":>6s} {"TB/s":>5s} {"floor us":>8s} {"floor/us":>8s} bound')
tot = tot_floor = 0.0
for k, (label, kernel, fl, by) in enumerate(launches, start=1):
    assert names[k].startswith(kernel), (k, label, names[k])
    d = dur[k]
    t_floor, m_floor = 2 * fl / TF * 1e6, by / HBM * 1e6
    floor = max(t_floor, m_floor)
    tot += d
    tot_floor += floor
    print(f'{label:46s} {kernel:>8s} {d:7.1f} {fl / 1e9:7.1f} {fl / d / 1e6:8.0f} {by / 1e6:6.0f} {by / d / 1e6:5.2f} {floor:8.1f} '
          f'{floor / d:8.2f} {"tensor" if t_floor >= m_floor else "hbm"}')
print(f'{len(launches)} trunk{" + FPN" if mode == "r02" else ""} launches: {tot:.0f} us measured, {tot_floor:.0f} us floor ({100 * tot_floor / tot:.0f} %); '
      f'peaks: {HBM / 1e9:.0f} GB/s, {TF / 1e12:.0f} TFLOP/s at 2 MMA units per algorithmic MMA')
