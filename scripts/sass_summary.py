"""Development aid: static SASS summary of the hot kernels of the in-tree library (what the ncu summaries cannot show
without a GPU: that the prefetches, cluster barriers and acquire loads are there, how much spill code is left).
usage: python scripts/sass_summary.py > profiles/r1_sass_summary.txt"""
import collections, os, re, subprocess, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wflow_b200", "libwflowsbm.so")
KERNELS = [
    ("k_vertical<4, DIAG=0, LAI=0>  (column, production build)", "_ZN2wf10k_verticalILi4ELb0ELb0EEEvNS_8DevModelE"),
    ("k_vertical<4, DIAG=1, LAI=1>  (column, every option)", "_ZN2wf10k_verticalILi4ELb1ELb1EEEvNS_8DevModelE"),
    ("k_lateral_wave<4, wide cluster, DIAG=0>  (sweep of the bench shard)", "_ZN2wf14k_lateral_waveILi4ELi3ELb0EEEvNS_8DevModelE"),
    ("k_lateral_wave<4, cluster, DIAG=0>", "_ZN2wf14k_lateral_waveILi4ELi1ELb0EEEvNS_8DevModelE"),
    ("k_lateral_wave<4, single block, DIAG=0>", "_ZN2wf14k_lateral_waveILi4ELi2ELb0EEEvNS_8DevModelE"),
    ("k_lateral_wave<4, grid, DIAG=0>", "_ZN2wf14k_lateral_waveILi4ELi0ELb0EEEvNS_8DevModelE"),
]
WATCH = ["CCTL", "UCGABAR_ARV", "UCGABAR_WAIT", "MEMBAR", "LDL", "STL", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "DMUL",
         "MUFU", "CALL", "BAR", "ATOM", "RED", "ERRBAR"]

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.startswith("api.") and f.endswith(".cubin")][0]
    print("# static SASS of %s (sm_100a), cuobjdump -sass -fun <kernel>; counts are instructions in the kernel's own text" % os.path.relpath(LIB, ROOT))
    print("# (the shared float64 routines wf_exp / wf_log / wf_div / unsatzone_flow_layer / act_transp_unsat / kinematic_wave* are separate functions)")
    for title, sym in KERNELS:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, cubin], capture_output=True, text=True).stdout
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", out):
            ops[m.group(1)] += 1
            if m.group(1) in ("CCTL", "LDG", "MEMBAR"):
                ops[m.group(1) + m.group(2)] += 1
        total = sum(v for k, v in ops.items() if "." not in k)
        print("\n%s\n  %d instructions (%d KB)" % (title, total, total * 16 // 1024))
        print("  " + "  ".join("%s %d" % (k, ops[k]) for k in WATCH if ops.get(k)))
        print("  " + "  ".join("%s %d" % (k, v) for k, v in sorted(ops.items()) if "." in k))
