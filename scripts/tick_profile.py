"""Development aid: summarise a WF_TICK_PROFILE csv (ns since the start of each tick)."""
import sys
import numpy as np
for path in sys.argv[1:]:
    a = np.genfromtxt(path, delimiter=",", names=True)
    sel = slice(len(a) // 10, len(a) * 9 // 10)
    print(path)
    for k in a.dtype.names[1:15]:
        print("  %-14s median %8.0f  p90 %8.0f max %8.0f" % (k, np.median(a[k][sel]), np.percentile(a[k][sel], 90), a[k][sel].max()))
    print("  tick period", np.median(np.diff(a["start_ns"][sel])))
