"""The CUDA path against trajectories of the reference's whole, unmodified ``WflowModel.dynamic()`` (PCRaster half
included: tests/golden/dyn_*.npz, written by oracle/gen_golden_dynamic.py through the float32 shim).

Kernel parity per step: every step starts from the reference's own float32 states of the step before, and every
state and requested output of every cell must agree within rel 1e-5 (absolute floor 2e-5 of the field's unit,
1e3 mm3 for subsurface flow; cancellation residuals by their own bound)."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

ATOL = {"SubsurfaceFlow": 1e3, "SoilWatbal": 1e-3, "watbal": 1e-3, "SurfaceWatbal": 1e-4, "InterceptionWatBal": 1e-4,
        "CellInFlow": 1e3}


@pytest.mark.parametrize("name", common.golden_dyn_cases())
def test_cuda_matches_reference_dynamic(name):
    g, kw, _, mod = common.load_dyn_case(name, oracle=False, gpu=True)
    try:
        nodes = mod.network()[0]
        mask = np.zeros(mod.shape, bool)
        mask.flat[nodes] = True
        from wflow_b200 import _lib
        known = set(_lib.field_names())
        skip = {"Cmax", "CanopyGapFraction", "EoverR"}  # inputs here; the reference overwrites them from LAI
        outs = [f for f in common.dyn_outputs(g) if f in known and f not in skip]
        states = [f for f in outs if "state0/" + f in g.files]
        mod.request(*outs)
        worst = (0.0, "")
        for t in range(kw["steps"]):
            for f in ("P", "PET", "TEMP", "Inflow"):
                key = "forcing/%d/%s" % (t, f)
                if key in g.files:
                    mod.set(f, g[key])
            for f in states:  # the reference's states of the step before, bit for bit
                mod.set(f, g["state0/" + f] if t == 0 else g["out/%d/%s" % (t - 1, f)])
            if kw.get("courant"):
                itl, itr = mod.estimate_iterations("land"), mod.estimate_iterations("river")
                assert (itl, itr) == tuple(int(v) for v in g["it/%d" % t]), (name, t)
                mod.set_substeps(itl, itr)
            mod.step()
            for f in outs:
                e = common.compare(mod.get(f), g["out/%d/%s" % (t, f)], mask, 1e-5, ATOL.get(f, 2e-5))
                if e > worst[0]:
                    worst = (e, "%s step %d" % (f, t))
                assert e <= 1.0, (name, t, f, e)
        print("%s: %d outputs x %d steps, worst error/tolerance %.3g (%s)" % (name, len(outs), kw["steps"], worst[0], worst[1]))
    finally:
        mod.close()
