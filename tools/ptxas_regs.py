#!/usr/bin/env python
"""Registers / spills per kernel from a .ptxas.log (names demangled); optional substring filters."""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                  r"(\d+) bytes spill loads\n.*?Used (\d+) registers", log)
dem = subprocess.run(['c++filt'] + [e[0] for e in ents], capture_output=True, text=True).stdout.split('\n')
for d, e in zip(dem, ents):
    if len(sys.argv) < 3 or any(f in d for f in sys.argv[2:]):
        print('%-100s stack %4s spill %4s/%-4s regs %s' % (d[:100], e[1], e[2], e[3], e[4]))
