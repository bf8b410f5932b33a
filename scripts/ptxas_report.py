"""Development aid: registers / spills per kernel from the nvcc -Xptxas -v log.  usage: make -C wflow_b200/csrc -B > log 2>&1; python scripts/ptxas_report.py log [filter]"""
import re
import subprocess
import sys

t = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else "(int)4"
pat = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\nptxas info\s+: Function properties for [^\n]+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
for name, st, ss, sl, regs in pat.findall(t):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if flt in d:
        print("%3s regs  stack %4s  spill st %4s ld %4s  %s" % (regs, st, ss, sl, d[:100]))
