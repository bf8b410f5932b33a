"""Development aid: turn the ncu reports a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/make_profiles.py [tag]     (tag of scripts/r2_gpu_run.sh, default r2)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def details(rep, kernel_substr, dest, header):
    text = ncu("-i", rep, "--page", "details")
    blocks, cur = [], []
    for line in text.splitlines():
        if line.startswith("  ") and not line.startswith("    ") and "(" in line and "Context" in line:
            if cur:
                blocks.append(cur)
            cur = [line]
        elif cur:
            cur.append(line)
    if cur:
        blocks.append(cur)
    pick = [b for b in blocks if kernel_substr in b[0]]
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    extra = []
    if len(raw) > 2:
        h = raw[0]
        for v in raw[2:]:
            d = dict(zip(h, v))
            if kernel_substr in d.get("Kernel Name", ""):
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                extra.append("# raw: gpu__time_duration.sum %s ms, dram__bytes_read.sum %s %s, dram__bytes_write.sum %s %s (sum %.3f), "
                             "smsp__inst_executed.sum %s, launch__registers_per_thread %s"
                             % (d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], raw[1][h.index("dram__bytes_read.sum")],
                                d["dram__bytes_write.sum"], raw[1][h.index("dram__bytes_write.sum")], rd + wr,
                                d["smsp__inst_executed.sum"], d["launch__registers_per_thread"]))
    with open(os.path.join(OUT, dest), "w") as fh:
        fh.write(header.rstrip() + "\n")
        for e in extra:
            fh.write(e + "\n")
        for b in pick[:1]:
            fh.write("\n".join(b) + "\n")
    print("wrote", dest, "(%d launches matched)" % len(pick))


def launches(src, dest, header):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    h = rows[0]
    ki, bi, gi, vi = h.index("Kernel Name"), h.index("Block Size"), h.index("Grid Size"), h.index("Metric Value")
    ui = h.index("Metric Unit")
    with open(os.path.join(OUT, dest), "w") as fh:
        fh.write(header.rstrip() + "\n")
        fh.write("id,kernel,block,grid,duration_us\n")
        for r in rows[1:]:
            v = float(r[vi].replace(",", ""))
            us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
            fh.write("%s,%s,%s,%s,%.1f\n" % (r[0], r[ki].split("(")[0].replace(", ", ";"), r[bi].replace(",", " "), r[gi].replace(",", " "), us))
    print("wrote", dest)


def traffic(rep, ncell, dest, source):
    """DRAM bytes per cell and launch of the two hot kernels -> profiles/<dest> (read by bench.py's roofline.traffic)."""
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    h, units = raw[0], raw[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    out = {}
    for v in raw[2:]:
        d = dict(zip(h, v))
        for key in ("k_vertical", "k_lateral_wave"):
            if key in d.get("Kernel Name", "") and key not in out:
                b = sum(float(d[c]) * scale[units[h.index(c)]] for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                out[key] = {"bytes_per_cell": b / ncell, "dram_bytes": b, "ms": float(d["gpu__time_duration.sum"]),
                            "inst_executed": float(d["smsp__inst_executed.sum"]), "kernel": d["Kernel Name"], "source": source}
    json.dump(out, open(os.path.join(OUT, dest), "w"), indent=1)
    print("wrote", dest, {k: round(v["bytes_per_cell"], 1) for k, v in out.items()})


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
    cmd = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-basin"
    note = ("# ncu --set full --clock-control none --import-source on, one launch of %s (round 2, default bench workload: "
            "32 basins of 1024^2 = 33.5 M cells)\n# command: " + cmd + " ; report read with: ncu -i prof_" + tag + ".ncu-rep --page details")
    details(rep, "k_vertical", "r2_vertical_ncu_details.txt", note % "k_vertical")
    details(rep, "k_lateral_wave", "r2_lateral_ncu_details.txt", note % "k_lateral_wave")
    traffic(rep, 33554431, "r2_traffic.json",
            "profiles/r2_*_ncu_details.txt (ncu --set full capture of the default workload, dram__bytes_read.sum + dram__bytes_write.sum per launch; static)")
    launches(os.path.join(G, "launches_%s.csv" % tag), "r2_launches_ncu.csv",
             "# ncu launch list (round 2): %s  (32 basins of 1024^2, default workload)\n"
             "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400; per-launch times are cold-cache and serialised: compare SHARES\n"
             "# (k_gather = uploads of the model's rasters and of the forcing; k_vertical + k_lateral_wave = one timestep)" % cmd)
