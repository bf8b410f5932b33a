"""Development aid: registers / stack / spills per kernel from an nvcc -Xptxas -v log.
usage: make -C wflow_b200/csrc dev > log 2>&1; python scripts/ptxas_report.py log [substring ...]"""
import re
import subprocess
import sys

t = open(sys.argv[1]).read()
want = sys.argv[2:] or ["k_vertical", "k_lateral", "k_pipeline"]
for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'(.*?)ptxas info\s+: Used (\d+) registers", t, re.S):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    fp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", m.group(2))
    if any(w in name for w in want):
        print("%3s regs  stack %4s  spill st %4s ld %4s  %s" % ((m.group(3),) + fp.groups() + (name[:100],)))
