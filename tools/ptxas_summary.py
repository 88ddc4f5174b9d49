#!/usr/bin/env python
"""Summarise registers / spills per kernel from the ptxas logs written by the Makefile (build/obj/*.ptxas.log)."""
import glob, re, subprocess, sys
pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'.*?Function properties for \1\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?", re.S)
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob('build/obj/*.ptxas.log')):
    txt = open(f).read()
    for m in pat.finditer(txt):
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void pirb::', '').replace('pirb::', '')
        if flt in name:
            print(f"{name:40s} regs={m.group(5):>4s} stack={m.group(2):>4s} spill_st={m.group(3):>4s} spill_ld={m.group(4):>4s}")
