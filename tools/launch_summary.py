"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family, share of the total.
    python tools/launch_summary.py gpurun_out/launches.csv [first_launch last_launch]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(unit, 1e-6)
        rows.append((int(r['ID']), r['Kernel Name'], v))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = [r for r in rows if lo <= r[0] <= hi]


def family(n):
    n = n.replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '')
    m = re.match(r'(flrelu_\w+::kernel|kernel)<([^>]*)>', n)
    if m:
        a = [s.strip() for s in m.group(2).split(',')]
        if len(a) == 5:
            return f'flrelu_stream::kernel<up{a[1]},fd{a[2]},signs{a[3]}>'
        return f'flrelu_bwd_stream::kernel<down{a[1]},signs{a[2]}>'
    n = re.sub(r'<.*', '', n)
    return n[:70]


tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, n, v in rows:
    a = agg[family(n)]
    a[0] += 1
    a[1] += v
print(f'{len(rows)} launches, {tot:.2f} ms in total (cold-cache, serialised: shares are meaningful, absolutes are not)')
print('| kernel | launches | ms | share |')
print('|---|---|---|---|')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f'| `{k}` | {c} | {v:.2f} | {100 * v / tot:.1f} % |')
