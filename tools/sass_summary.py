"""Per-kernel SASS evidence from the built library (no GPU needed): python tools/sass_summary.py [> profiles/rNN_sass_summary.md]

`cuobjdump -sass stylegan3-editing_b200/libsg3_b200.so`, per kernel the counts of the mnemonics that prove what the code is:
UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32), FFMA, FMNMX3, HMMA (mma.sync -- none expected), LDS / STS, LDL / STL (spills).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'stylegan3-editing_b200', 'libsg3_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FMNMX3',
       'HMMA', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL']

sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda names: subprocess.run(['cu++filt'] + names, stdout=subprocess.PIPE, text=True).stdout.split('\n')
kernels, cur = collections.OrderedDict(), None
for line in sass.split('\n'):
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        kernels[cur]['_total'] += 1
names = list(kernels)
pretty = dict(zip(names, demangle(names)))
sys.path.insert(0, ROOT)
import sg3_b200  # noqa: E402
print('# SASS summary of libsg3_b200.so')
print()
print(f'`{sg3_b200.capi.lib().sg3_build_info().decode()}`; `cuobjdump -sass`, instruction counts per kernel (static).  '
      f'{len(kernels)} kernels.')
print()
tot = collections.Counter()
for k in kernels.values():
    tot.update(k)
print('Whole library: ' + ', '.join(f'{op} {tot[op]}' for op in OPS if tot[op]))
print()
print('| kernel | instr | ' + ' | '.join(OPS) + ' |')
print('|---|---|' + '---|' * len(OPS))


def short(n):
    n = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', n)
    n = re.sub(r'\(flrelu_\w+::Params\)|\([^)]*\)$', '', n)
    n = n.replace('void ', '').replace('(int)', '').replace('(bool)', '')
    return n[:90]


groups = collections.OrderedDict()
for n, c in kernels.items():
    key = re.sub(r'<.*', '', short(pretty[n]))
    groups.setdefault(key, []).append((short(pretty[n]), c))
for key, items in groups.items():
    if len(items) > 6:      # template families: one summed row + the heaviest instance
        s = collections.Counter()
        for _, c in items:
            s.update(c)
        print(f'| `{key}<...>` ({len(items)} instances, summed) | {s["_total"]} | ' + ' | '.join(str(s[o]) if s[o] else '' for o in OPS) + ' |')
        nm, c = max(items, key=lambda it: it[1]['_total'])
        print(f'| `{nm}` (largest) | {c["_total"]} | ' + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
    else:
        for nm, c in items:
            print(f'| `{nm}` | {c["_total"]} | ' + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
