#!/usr/bin/env python
"""Per-kernel-name launches per step and mean / total duration of a tools/step_timeline.py CSV."""
import csv
import sys
from collections import defaultdict

for f in sys.argv[1:]:
    rows = list(csv.DictReader(open(f)))
    # one Adam launch per step (the fused tail removed the per-step weight split that used to mark a step)
    starts = [i for i, r in enumerate(rows) if 'adam_multi' in r['name']]
    if len(starts) < 3:
        print(f, 'too few steps')
        continue
    a, b = starts[1], starts[-1]
    nsteps = len(starts) - 2
    span = (float(rows[b]['start_us']) - float(rows[a]['start_us'])) / nsteps
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[a:b]:
        k = r['name'].split('(')[0].replace('void ', '')[:60]
        agg[k][0] += 1
        agg[k][1] += float(r['dur_us'])
    print('== %s: %.1f us/step, %.1f kernels/step, busy %.1f us/step' % (f, span, sum(v[0] for v in agg.values()) / nsteps,
                                                                        sum(v[1] for v in agg.values()) / nsteps))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('   %-62s x%4.1f  mean %6.2f us  total/step %6.2f' % (k, v[0] / nsteps, v[1] / v[0], v[1] / nsteps))
