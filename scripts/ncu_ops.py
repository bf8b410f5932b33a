"""Development aid: aggregate an `ncu --page source --print-source cuda,sass --csv` export by SASS opcode and by stall reason.
usage: python scripts/ncu_ops.py x.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
ops = defaultdict(lambda: [0.0, 0.0, 0.0])
stalls = defaultdict(float)
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ci = {}
        for i, k in enumerate(hdr):
            ci.setdefault(k, i)
        sass_i = [i for i, k in enumerate(hdr) if k == "Source"][1]
        stall_cols = [(k, i) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    s = r[sass_i].strip()
    if not s or s == "-" or not r[ci["Address"]].strip().startswith("0x"):
        continue
    tok = s.split()
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    op = op.rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("IMAD", "LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU") else op.split(".")[0]
    try:
        n = float(r[ci["Instructions Executed"]]); th = float(r[ci["Thread Instructions Executed"]]); sm = float(r[ci["# Samples"]])
    except ValueError:
        continue
    ops[base][0] += n; ops[base][1] += th; ops[base][2] += sm
    for k, i in stall_cols:
        try:
            stalls[k] += float(r[i])
        except ValueError:
            pass
tot = sum(v[0] for v in ops.values()) or 1
tots = sum(v[2] for v in ops.values()) or 1
print("warp instructions %.4g, samples %d" % (tot, tots))
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %5.1f%% inst  %5.1f%% samples  %4.1f lanes" % (k, 100 * v[0] / tot, 100 * v[2] / tots, v[1] / max(v[0], 1)))
st = sum(stalls.values()) or 1
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / st) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:12]))
