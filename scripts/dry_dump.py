"""Development aid: GPU kinematic_wave_ssf on the dry-column golden set, saved for offline comparison."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wflow_b200 import core
g = np.load("tests/golden/kinwave_ssf_dry.npz")
np.save("gpurun_out/dry_out.npy", core.kinematic_wave_ssf(*g["args"]))
