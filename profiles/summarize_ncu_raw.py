#!/usr/bin/env python
"""Key metrics per captured launch from an `ncu --set full` report exported with
    ncu -i <report>.ncu-rep --page raw --csv > <name>.raw.csv
(tools/gpu_evidence.sh does the export on the GPU box: the reports themselves are too large to bring back).

    python profiles/summarize_ncu_raw.py <name>.raw.csv "<how it was captured>" [kernel-substring] > profiles/<name>.json

bench.py reads `traffic_bytes_per_launch` of the SpMM summaries for `roofline.traffic`."""
import csv
import json
import sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
TIME = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return None


def main(path, source, only=''):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, k, table=None):
        if k not in col:
            return None
        v = num(r[col[k]])
        if v is None:
            return None
        u = units[col[k]]
        if table is not None:
            return v * table.get(u, 1)
        return v
    out = []
    for r in body:
        name = r[col['Kernel Name']]
        if only and only not in name:
            continue
        rd, wr = g(r, 'dram__bytes_read.sum', UNIT), g(r, 'dram__bytes_write.sum', UNIT)
        us = g(r, 'gpu__time_duration.sum', TIME)
        e = {
            'kernel': name.split('(')[0].replace('void ', '').replace('gist::', ''),
            'grid': r[col['Grid Size']], 'block': r[col['Block Size']], 'us': us,
            'regs': g(r, 'launch__registers_per_thread'),
            'dram_read_bytes': rd, 'dram_write_bytes': wr,
            'dram_GBps': round((rd + wr) / us / 1e3, 1) if us else None,
            'l2_hit_pct': g(r, 'lts__t_sector_hit_rate.pct'),
            'xbar_to_l1_bytes': g(r, 'l1tex__m_xbar2l1tex_read_bytes.sum', UNIT),
            'tma_load_bytes': g(r, 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', UNIT),
            'warps_active_per_sm': g(r, 'sm__warps_active.avg.per_cycle_active'),
            'sm_throughput_pct': g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
            'issue_slots_busy_pct': g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            'tensor_pipe_active_pct': g(r, 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
            'tensor_inst_pct_of_peak': g(r, 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active'),
        }
        if e['xbar_to_l1_bytes'] and us:
            e['xbar_to_l1_TBps'] = round(e['xbar_to_l1_bytes'] / us / 1e6, 2)
        out.append({k: v for k, v in e.items() if v is not None})
    tot = sum(l.get('dram_read_bytes', 0) + l.get('dram_write_bytes', 0) for l in out)
    print(json.dumps({'source': source, 'traffic_bytes_per_launch': round(tot / max(len(out), 1)), 'launches': out}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '', sys.argv[3] if len(sys.argv) > 3 else '')
