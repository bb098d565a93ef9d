#!/usr/bin/env python
"""Per-launch DRAM traffic of the SpMM kernels from an `ncu --set full` report.

    ncu -i <report>.ncu-rep --page raw --csv > raw.csv
    python profiles/extract_traffic.py raw.csv "<how the report was captured>" > profiles/<name>.json

bench.py reads `traffic_bytes_per_launch` (mean of dram__bytes_read.sum + dram__bytes_write.sum
over the captured launches) for `roofline.traffic`."""
import csv
import json
import sys


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return None


def to_bytes(v, unit):
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main(path, source):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in body:
        name = r[col['Kernel Name']]
        if 'spmm' not in name:
            continue
        g = lambda k: (num(r[col[k]]), units[col[k]])          # noqa: E731
        rd, wr = g('dram__bytes_read.sum'), g('dram__bytes_write.sum')
        dur = g('gpu__time_duration.sum')
        out.append({
            'kernel': name.split('(')[0].replace('void ', '').replace('gist::', ''),
            'us': dur[0] * {'us': 1, 'ms': 1e3, 'ns': 1e-3}[dur[1]],
            'dram_read_bytes': to_bytes(*rd), 'dram_write_bytes': to_bytes(*wr),
            'l2_hit_pct': num(r[col['lts__t_sector_hit_rate.pct']]),
            'xbar_to_l1_bytes': to_bytes(*g('l1tex__m_xbar2l1tex_read_bytes.sum')),
            'warps_active_per_sm': num(r[col['sm__warps_active.avg.per_cycle_active']]),
            'grid': r[col['Grid Size']], 'regs': num(r[col['launch__registers_per_thread']]),
        })
    tot = sum(l['dram_read_bytes'] + l['dram_write_bytes'] for l in out)
    print(json.dumps({'source': source, 'traffic_bytes_per_launch': round(tot / max(len(out), 1)),
                      'launches': out}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
