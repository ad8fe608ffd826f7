#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_ncu.py launches <launches.csv> <out.md>
    python tools/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    iN, iV = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in data:
        n = r[iN].split('(')[0].replace('svb::', '').replace('void ', '')
        v = float(r[iV].replace(',', '')) / 1000.0
        tot[n] = tot.get(n, 0) + v
        cnt[n] += 1
    T = sum(tot.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list: {len(data)} launches, {T:.1f} us total (serialised, cold cache: compare SHARES)\n\n')
        f.write('| kernel | launches | total us | share |\n|---|---:|---:|---:|\n')
        for n, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f'| `{n}` | {cnt[n]} | {v:.1f} | {100 * v / T:.1f} % |\n')
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    iN = hdr.index('Kernel Name')
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary of {src}\n\n')
        f.write('| metric | unit | ' + ' | '.join(f'launch {i}' for i in range(len(data))) + ' |\n')
        f.write('|---|---|' + '---:|' * len(data) + '\n')
        f.write('| kernel | | ' + ' | '.join(d[iN].split('(')[0][-28:] for d in data) + ' |\n')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f'| `{k}` | {units[i]} | ' + ' | '.join(d[i] for d in data) + ' |\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
