"""Run the reference's shipped Rhine SBM case (a local, git-ignored copy under baseline/_ref/) through
the framework mirror and print the number the reference's own test pins
(tests/test_wflow_sbm_example.py:54: sum of specrun.csv column 2 over steps 1..29 = 0.0432277)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wflow_b200.framework import WflowSbm

case = sys.argv[1] if len(sys.argv) > 1 else "baseline/_ref/wflow_rhine_sbm"
m = WflowSbm(Dir=case, RunDir="unittestsbm", configfile="wflow_sbm.ini")
m.createRunId()
t0 = time.time(); m._runInitial(); m._runResume(); t1 = time.time()
for ts in range(1, 30):
    m._runDynamic(ts, ts)
t2 = time.time()
cells = m._mod.num_cells
print("cells", cells, "levels", m._mod.num_levels, "river cells", m._mod.num_river_cells)
print("init %.2f s, 29 steps %.3f s (incl. map-stack reads and csv/tss/map output)" % (t1 - t0, t2 - t1))
m._runSuspend(); m._wf_shutdown()
d = np.genfromtxt(os.path.join(case, "unittestsbm", "specrun.csv"), delimiter=",")
print("specrun.csv shape", d.shape, "sum column 2 = %.9f (reference test expects 0.043227685 +- 5e-5)" % d[:, 2].sum())
p = np.genfromtxt(os.path.join(case, "unittestsbm", "prec.csv"), delimiter=",")
print("prec.csv sum column 2 = %.6f" % p[:, 2].sum())
