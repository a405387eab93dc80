"""CPU-only: per-kernel histogram of the Blackwell-specific SASS mnemonics in libmcgaze_b200.so (cuobjdump -sass), as a
markdown table: tcgen05.mma (UTCHMMA fp16, UTCQMMA fp8), tcgen05.ld (LDTM), tcgen05.commit (UTCBAR), TMA loads / stores
(UTMALDG / UTMASTG), bulk copies (UBLKCP), legacy mma.sync (HMMA).  usage: python tools/sass_histogram.py [lib] > profiles/...md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'mcgaze_b200', 'libmcgaze_b200.so')
MN = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'HMMA', 'SYNCS', 'LDGSTS']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m:
        op = m.group(1)
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
                break
        counts[cur]['_total'] += 1
print(f'SASS instruction histogram of `{os.path.relpath(lib, ROOT)}` (sm_100a), static counts per kernel\n')
print('| kernel | instructions | ' + ' | '.join(MN) + ' |')
print('|---|---|' + '---|' * len(MN))
tot = collections.Counter()
for fn, c in counts.items():
    name = re.sub(r'\(.*', '', demangle(fn)).replace('void ', '').replace('mcg::', '')
    print(f'| `{name}` | {c["_total"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in MN) + ' |')
    tot.update(c)
print(f'| **total** | {tot["_total"]} | ' + ' | '.join(str(tot[k]) for k in MN) + ' |')
