"""Summarise one .ncu-rep capture of a stream kernel: duration, pipe / issue / shared-memory utilisation, DRAM bytes,
dynamic SASS opcode mix (per warp-level group iteration when --groups is given).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--groups N]      (N = number of warp group iterations of the launch)
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
groups = float(sys.argv[sys.argv.index('--groups') + 1]) if '--groups' in sys.argv else None


def run(args):
    return subprocess.run(['ncu', '-i', rep] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


raw = list(csv.reader(io.StringIO(run(['--page', 'raw', '--csv']))))
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active_mem_lgds.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct']
for k in keys:
    for h in m:
        if h == k:
            print(f'{k:85s} {m[h][0]:>16s} {m[h][1]}')
print()
for h in m:
    if 'pipe' in h and 'pct_of_peak_sustained_active' in h and 'inst_executed' not in h:
        try:
            if float(m[h][0]) > 5:
                print(f'  {h:80s} {m[h][0]}')
        except ValueError:
            pass
src = list(csv.reader(io.StringIO(run(['--page', 'source', '--csv']))))
h2 = src[1]
iS, iE, iW, iSm = h2.index('Source'), h2.index('Instructions Executed'), h2.index('L1 Wavefronts Shared'), h2.index('# Samples')
byop, wf, smp = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in src[2:]:
    toks = [t for t in r[iS].split() if not t.startswith('@')]
    op = toks[0].split('.')[0]
    n = int(r[iE])
    tot += n
    byop[op] += n
    wf[op] += int(r[iW] or 0)
    smp[op] += int(r[iSm] or 0)
print(f'\nwarp instructions executed: {tot}' + (f' = {tot / groups:.1f} per group' if groups else ''))
for op, n in byop.most_common(26):
    line = f'  {op:10s} {100 * n / tot:5.1f}%'
    if groups:
        line += f'  {n / groups:7.1f}/group'
    if wf[op]:
        line += f'   smem wavefronts {wf[op] / groups if groups else wf[op]:.1f}' + ('/group' if groups else '')
    line += f'   stall samples {smp[op]}'
    print(line)
if groups:
    print(f'  smem wavefronts total {sum(wf.values()) / groups:.1f}/group')
st = collections.Counter()
for h in m:
    if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
        try:
            st[h[len('smsp__pcsamp_warps_issue_stalled_'):]] += int(float(m[h][0]))
        except ValueError:
            pass
print('\nstall samples:', ', '.join(f'{k} {v}' for k, v in st.most_common(8)))
