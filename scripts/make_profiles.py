"""Development aid: turn the ncu reports a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.
usage: python scripts/make_profiles.py"""
import csv
import io
import os
import subprocess

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


if __name__ == "__main__":
    rep = os.path.join(G, "prof_r1_final.ncu-rep")
    cmd = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-basin"
    details(rep, "k_vertical", "r1_vertical_ncu_details.txt",
            "# ncu --set full --clock-control none --import-source on, one launch of k_vertical (round 1, default bench workload: "
            "32 basins of 1024^2 = 33.5 M cells)\n# command: %s ; report read with: ncu -i prof_r1_final.ncu-rep --page details" % cmd)
    details(rep, "k_lateral_wave", "r1_lateral_ncu_details.txt",
            "# ncu --set full --clock-control none --import-source on, one launch of k_lateral_wave (round 1, default bench workload: "
            "32 basins of 1024^2 = 33.5 M cells)\n# command: %s ; report read with: ncu -i prof_r1_final.ncu-rep --page details" % cmd)
    pipe = os.path.join(G, "prof_pipe_full2.ncu-rep")
    if os.path.isfile(pipe):
        details(pipe, "k_pipeline", "r1_pipeline_ncu_details.txt",
                "# ncu --set full --clock-control none --import-source on, one launch of k_pipeline (round 1): 20 timesteps in one launch, "
                "the default bench shard (32 basins of 1024^2 = 33.5 M cells) swept as one group by the cooperative grid kernel,\n"
                "# chunks dealt on demand inside a block; development build with 1024 threads per block (64 registers; the default is 768 / 80)\n"
                "# command: python scripts/pipe_bench.py --basins 32 --modes grid+lib:libwflowsbm_pfh.so --ks 20")
    old = os.path.join(G, "prof_pipe_grid.ncu-rep")
    if os.path.isfile(old):
        details(old, "k_pipeline", "r1_pipeline_static_ncu_details.txt",
                "# ncu --set full, one launch of k_pipeline BEFORE chunks were dealt on demand (fixed item -> thread map, 3 blocks of 256 threads "
                "per SM): 8 timesteps, 8 basins of 1024^2; 57 % of the stall cycles at the block barrier ahead of the grid barrier\n"
                "# command: python scripts/pipe_bench.py --basins 8 --modes grid --ks 8")
    launches(os.path.join(G, "launches_r1.csv"), "r1_launches_ncu.csv",
             "# ncu launch list (round 1): %s  (32 basins of 1024^2, default workload)\n"
             "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400; per-launch times are cold-cache and serialised: compare SHARES\n"
             "# (k_gather = uploads of the model's rasters and of the forcing; k_vertical<4;0> + k_lateral_wave<4;3;0> = one timestep:\n"
             "#  column 4.64-4.73 ms + sweep 7.90-8.05 ms here, 5.02 + 7.88 ms by CUDA events inside bench.py -- share of the step 37 %% / 63 %% vs 39 %% / 61 %%)" % cmd)
