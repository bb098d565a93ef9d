#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: summarize_launches.py launches.csv [steps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, tot, n = collections.OrderedDict(), 0.0, 0
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    name = re.sub(r'\(.*', '', r[ki])[:100]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    n += 1
print('total %.1f us over %d launches (%.1f us, %.1f launches per step)' % (tot, n, tot / steps, n / steps))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%5.1f%% %8.1f us/step  n/step=%5.1f  %s' % (100 * t / tot, t / steps, c / steps, k))
