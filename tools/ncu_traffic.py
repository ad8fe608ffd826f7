#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernel family (merged ResBlock steps of conv1d_c4_tc_kernel) from an
`ncu --set full` capture of ONE forward (tools/ncu_forward.py, -s 58 -c 29).
    python tools/ncu_traffic.py gpurun_out/prof_fwd.ncu-rep profiles/r02_conv_tc_traffic.json [profiles/r02_conv_tc_ncu_full.md]"""
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
    rep, out_json = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in KEYS if k in hdr}
    launches = []
    for i, d in enumerate(data):
        rd = to_bytes(d[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
        wr = to_bytes(d[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
        # order inside a forward: 0 conv_pre; then per stage: upsampler, 6 merged ResBlock steps
        kind = 'conv_pre' if i == 0 else ('upsampler' if (i - 1) % 7 == 0 else 'resblock step (3 layers)')
        launches.append({'i': i, 'kind': kind, 'us': float(d[col['gpu__time_duration.sum']].replace(',', '')) / (1e3 if units[col['gpu__time_duration.sum']] == 'ns' else 1),
                         'dram_read': rd, 'dram_write': wr,
                         **{k.split('.')[0]: d[c] for k, c in col.items() if k not in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum')}})
    fam = [x for x in launches if x['kind'].startswith('resblock')]
    res = {'source': rep, 'launches_in_family': len(fam), 'mean_bytes_per_launch': sum(x['dram_read'] + x['dram_write'] for x in fam) / max(1, len(fam)),
           'total_bytes_family': sum(x['dram_read'] + x['dram_write'] for x in fam),
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the merged ResBlock-step launches of one forward (ncu --set full, tools/ncu_forward.py)',
           'launches': launches}
    json.dump(res, open(out_json, 'w'), indent=1)
    if len(sys.argv) > 3:
        with open(sys.argv[3], 'w') as f:
            f.write(f'# ncu --set full of one generator forward (config 2, bf16x3), {len(launches)} launches of conv1d_c4_tc_kernel\n\n')
            f.write('| # | kind | us | dram read MB | dram write MB | tensor pipe % | l1tex % | lts % | dram % | sm % | warps active % |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
            for x in launches:
                f.write(f"| {x['i']} | {x['kind']} | {x['us']:.1f} | {x['dram_read'] / 1e6:.1f} | {x['dram_write'] / 1e6:.1f} | {x.get('sm__pipe_tensor_cycles_active', '')} | "
                        f"{x.get('l1tex__throughput', '')} | {x.get('lts__throughput', '')} | {x.get('gpu__dram_throughput', '')} | {x.get('sm__throughput', '')} | {x.get('sm__warps_active', '')} |\n")
    print(json.dumps({k: v for k, v in res.items() if k != 'launches'}, indent=1))


if __name__ == '__main__':
    main()
