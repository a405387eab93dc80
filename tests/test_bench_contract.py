"""CPU: the committed bench line (profiles/r01_bench_g.json, written by bench.py on a B200) carries every key of the
bench contract, and bench.py's static pieces (metric string, workload constants) agree with BASELINE.json."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_g.json')))
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['unit'] == 'clips/s' and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['warmup'] >= 3
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and r['traffic']
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 32 * 7 * 3 * 224 * 224 * 4 and e['d2h_bytes_per_step'] > 0 and e['value'] != d['value']
    c = d['cpu_baseline']
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['sample']
    assert d['gpu_launches'] == 147 * d['steps']
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    p = d['preprocess']
    assert p['kernel']['bound'] == 'hbm' and 0 < p['kernel']['frac'] < 1 and p['e2e_u8']['h2d_bytes_per_step'] < e['h2d_bytes_per_step']


def test_bench_measures_the_baseline_metric():
    src = open(os.path.join(ROOT, 'bench.py')).read()
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert 'clips/sec' in base['metric'] and re.search(r"METRIC = '(clips/sec[^']*)'", src)
    assert re.search(r'CLIPS_PER_STEP\s*=\s*32', src) and re.search(r'T, H, W = 7, 224, 224', src)
    assert "'--impl'" in src and "'reference'" in src
