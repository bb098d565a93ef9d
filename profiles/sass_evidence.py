#!/usr/bin/env python
"""Counts the SASS mnemonics that show what each kernel of libgist_b200.so is built from
(B200_PROFILING.md, "What proves a Blackwell-native kernel"):  python profiles/sass_evidence.py
> profiles/r1_sass_evidence.txt   (CPU only: cuobjdump on the cross-compiled library)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'gist_b200', 'csrc', 'libgist_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'HGMMA',
         'SYNCS', 'FADD2', 'FFMA2', 'IMAD.WIDE.U32', 'LDG.E.128', 'LDG.E.64', 'ATOMG', 'SHFL', 'MEMBAR', 'ERRBAR',
         'ACQBULK', 'PREEXIT']      # griddepcontrol.wait / .launch_dependents (programmatic dependent launch)


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r'\(.*$', '', kern).replace('void ', '').replace('(anonymous namespace)::', '')
            counts[kern] = collections.Counter()
            total[kern] = 0
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and kern:
            total[kern] += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    counts[kern][w] += 1
    print('# %s — SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)' % os.path.relpath(LIB, ROOT))
    print('# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA load, SYNCS = mbarrier ops,')
    print('# FADD2 / FFMA2 = packed fp32x2 arithmetic, HMMA would be the legacy mma.sync path (absent)')
    for k in sorted(counts):
        c = counts[k]
        if not any(c[w] for w in WATCH if w not in ('SHFL', 'LDG.E.128', 'LDG.E.64')):
            continue
        print('%-70s %5d instr  %s' % (k[:70], total[k], '  '.join('%s=%d' % (w, c[w]) for w in WATCH if c[w])))


if __name__ == '__main__':
    main()
