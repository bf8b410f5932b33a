"""Development aid: time k_vertical alone (all-pit LDD: no routing) for one or several builds of the library.
usage: python scripts/vert_bench.py [rows cols]         (uses WFLOW_B200_LIB or the in-tree library)
       python scripts/vert_bench.py compare A B C ...    (runs itself once per wflow_b200/libwflowsbm_<X>.so)"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "compare":
    for tag in sys.argv[2:]:
        lib = os.path.join(ROOT, "wflow_b200", "libwflowsbm%s.so" % ("" if tag == "-" else "_" + tag))
        env = dict(os.environ, WFLOW_B200_LIB=lib)
        out = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        print(tag, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
    sys.exit(0)

from wflow_b200 import synth
from wflow_b200.core import SbmModel
nr, nc = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 4096)
lt, dt = (100, 300, 800), 3600
cache = "/tmp/vert_case_%d_%d.npz" % (nr, nc)
if os.path.isfile(cache):
    z = np.load(cache)
    fields = {k: z[k] for k in z.files}
else:
    ldd, river, static, state = synth.make_case(nr, nc, kind="pits", timestepsecs=dt, layer_thickness=lt, init="random")
    fields = {k: v for k, v in {**static, **state}.items() if k not in ("River", "Beta", "SatWaterDepth")}
    np.savez(cache, **{k: np.broadcast_to(np.asarray(v, np.float32), (nr, nc)) for k, v in fields.items()})
ldd = np.full((nr, nc), 5, np.uint8)
mod = SbmModel(ldd, None, timestepsecs=dt, layer_thickness=lt)
for k, v in fields.items():
    mod.set(k, v)
forc = synth.Forcing(nr, nc, timestepsecs=dt)
vs = []
for s in range(24):
    if s % 6 == 0:
        P, PET, T = forc.step(s // 6)
        mod.set("P", P); mod.set("PET", PET); mod.set("TEMP", T)
    mod.step()
    vs.append(mod.last_step_ms()[0])
v = float(np.median(vs[4:]))
n = mod.num_cells
chk = float(np.nansum(mod.get("zi").astype(np.float64)) + np.nansum(mod.get("UStoreLayerDepth_1").astype(np.float64)))
print("cells %d vertical %.3f ms (min %.3f) = %.2f Gcell/s, %.0f GB/s @236 B | checksum %.6f" % (n, v, min(vs[4:]), n / v / 1e6, n * 236 / v / 1e6, chk))
