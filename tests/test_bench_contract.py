"""CPU: bench.py builds every key of the bench contract (checked on its source), the newest committed bench line
(profiles/r*_bench_*.json, written by bench.py on a B200) carries them, and bench.py's static pieces (metric string,
workload constants) agree with BASELINE.json.  Keys and invariants only: no launch counts or timings are pinned."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONTRACT = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline')


def test_bench_source_builds_every_contract_key():
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for k in CONTRACT:
        assert re.search(r"['\"]%s['\"]\s*[:\]]" % k, src), k
    for k in ('bound', 'achieved', 'peak', 'frac', 'traffic', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'cores', 'kind',
              'sample', 'sm_mhz', 'sm_max_mhz', 'reasons', 'workload'):
        assert re.search(r"['\"]%s['\"]" % k, src), k


def test_committed_bench_line_has_the_contract_keys():
    lines = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_bench_*.json')))
    lines = [p for p in lines if 'baseline' not in os.path.basename(p)]
    assert lines, 'no committed bench line under profiles/'
    d = json.load(open(lines[-1]))
    for k in CONTRACT:
        assert k in d, k
    assert d['unit'] == 'clips/s' and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['warmup'] >= 3
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and r['traffic']
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 32 * 7 * 3 * 224 * 224 * 4 and e['d2h_bytes_per_step'] > 0 and e['value'] != d['value']
    c = d['cpu_baseline']
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['sample']
    assert d['gpu_launches'] > 0 and d['gpu_launches'] % d['steps'] == 0
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    p = d['preprocess']
    assert p['kernel']['bound'] == 'hbm' and 0 < p['kernel']['frac'] < 1 and p['e2e_u8']['h2d_bytes_per_step'] < e['h2d_bytes_per_step']


def test_bench_measures_the_baseline_metric():
    src = open(os.path.join(ROOT, 'bench.py')).read()
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert 'clips/sec' in base['metric'] and re.search(r"METRIC = '(clips/sec[^']*)'", src)
    assert re.search(r'CLIPS_PER_STEP\s*=\s*32', src) and re.search(r'T, H, W = 7, 224, 224', src)
    assert "'--impl'" in src and "'reference'" in src
