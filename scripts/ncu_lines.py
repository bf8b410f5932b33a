"""Development aid: aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.
usage: ncu -i rep --page source --print-source cuda,sass --csv > x.csv; python scripts/ncu_lines.py x.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ci = {k: i for i, k in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (fname, line)
    a = agg[key]
    a[3] = r[1]
    def num(name):
        try:
            return float(r[ci[name]])
        except Exception:
            return 0.0
    a[0] += num("# Samples")
    a[1] += num("Instructions Executed")
    a[2] += num("stall_no_inst")
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst %5.1f%% noinst  %s:%d  %s" % (100 * a[0] / tot_s, 100 * a[1] / tot_i, 100 * a[2] / max(a[0], 1), key[0], key[1], a[3].strip()[:90]))
