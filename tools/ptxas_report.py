#!/usr/bin/env python
"""Registers / spills of every chain and mixture kernel from build/ptxas.log."""
import re, subprocess, sys
log = open('build/ptxas.log').read()
ents = re.findall(r"Compiling entry function '([^']+)'.*?\n(?:.*\n)*?ptxas info\s+: Function properties for .*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", log)
dem = subprocess.run(['c++filt'] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
pat = sys.argv[1] if len(sys.argv) > 1 else 'MathFast'
for d, e in zip(dem, ents):
    d = d.replace('nfn::', '')
    if re.search(pat, d) and ('chain_kernel' in d or 'mdn_kernel' in d or 'kmn_kernel' in d):
        print('%3s regs  stack %4s  spill st/ld %4s/%4s | %s' % (e[4], e[1], e[2], e[3], d[:150]))
