"""CPU-only: per-layer roofline floors of the trunk against a committed ncu launch list.

    python tools/layer_floor.py [profiles/r01_launches_fp16c8_g.csv]

For each of the 52 trunk GEMM launches (ResNet-50 bottlenecks at 224^2, 32 clips x 7 frames) prints the measured
duration, the algorithmic FLOPs and bytes (unfused activations: 4 B / element in fp16c8 = fp16 hi + e4m3 lo8 + e4m3 hi8,
3 B for tensors only read as a residual), and the floor max(bytes / HBM peak, 2 x FLOPs / tensor peak) - fp16c8 spends
two fp16-MMA units per algorithmic MMA.  Peaks from MEASURED_PEAKS.json (fallback:
B200_PROFILING.md numbers)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r01_launches_fp16c8_g.csv')
try:
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    HBM, TF = pk['hbm_gbs'] * 1e9, pk['bf16_tflops_sustained'] * 1e12
except Exception:
    HBM, TF = 6650e9, 1400e12
rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
dur = [float(r['Metric Value'].replace(',', '')) / 1e3 for r in rows]          # us
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void mcg::', '') for r in rows]
assert 'stem_fused' in names[0], 'expects the launch list of exactly one step (tools/run_round.sh)'

NB, h, cin = 224, 56, 64
layers = []
for l, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
    for b in range(blocks):
        stride = 2 if (b == 0 and l > 0) else 1
        ho = h // stride
        # name, M, N, K, input pixels read, input channels, out B/elt, residual B/elt
        layers.append((f'layer{l + 1}.{b}.conv1', NB * h * h, planes, cin, NB * h * h, cin, 4, 0))
        layers.append((f'layer{l + 1}.{b}.conv2', NB * ho * ho, planes, 9 * planes, NB * h * h, planes, 4, 0))
        if b == 0:
            layers.append((f'layer{l + 1}.{b}.downsample', NB * ho * ho, 4 * planes, cin, NB * h * h, cin, 3, 0))
        layers.append((f'layer{l + 1}.{b}.conv3', NB * ho * ho, 4 * planes, planes, NB * ho * ho, planes, 4, 3))
        cin, h = 4 * planes, ho
print(f'{"layer":20s} {"kernel":>12s} {"us":>7s} {"GFLOP":>7s} {"TFLOP/s":>8s} {"MB":>6s} {"TB/s":>5s} {"floor us":>8s} {"floor/us":>8s}')
tot = tot_floor = 0.0
for k, (n, M, N, K, inpix, cin_, ob, rb) in enumerate(layers, start=1):
    d = dur[k]
    fl = 2.0 * M * N * K
    by = inpix * cin_ * 4 + M * N * (ob + rb) + N * K * 4
    floor = max(2 * fl / TF, by / HBM) * 1e6
    tot += d
    tot_floor += floor
    print(f'{n:20s} {names[k][-10:]:>12s} {d:7.1f} {fl / 1e9:7.1f} {fl / d / 1e6:8.0f} {by / 1e6:6.0f} {by / d / 1e6:5.2f} {floor:8.1f} {floor / d:8.2f}')
print(f'trunk GEMM launches: {tot:.0f} us measured, {tot_floor:.0f} us floor ({100 * tot_floor / tot:.0f} %)')
