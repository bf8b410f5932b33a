"""Live check of the C oracle against the reference's numba kernels (only where the reference
checkout is present, i.e. the build container; skipped on the GPU box)."""
import numpy as np
import pytest

from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference checkout not present")


def test_oracle_matches_reference_kernels_live():
    from oracle import gen_golden, ref_driver
    from wflow_b200 import synth
    kw = dict(n=18, kind="forest", layer_thickness=(120, 250), steps=6, seed=9, itL=2)
    orc, ldd, river, static, state = gen_golden.build_case(**kw)
    ref = ref_driver.ReferenceModel(orc)
    for a, b in zip(orc.net.nodes, ref.nodes):
        assert np.array_equal(a, b)
    for a, b in zip(orc.net.nodes_up, ref.nodes_up):
        assert np.array_equal(a, b)
    forc = synth.Forcing(kw["n"], kw["n"], seed=6)
    orc.want(*gen_golden.PHASE_DIAGS)
    outs = ["SubsurfaceFlow", "zi", "LandRunoff", "RiverRunoff", "WaterLevelR", "UStoreLayerDepth_0", "UStoreLayerDepth_2"]
    mask = np.zeros(orc.shape, bool)
    mask.flat[orc.net.flat_nodes] = True
    pre_names = ["zi", "SubsurfaceFlow", "WaterLevelL", "LandRunoffsub", "RiverRunoffsub"] + \
                ["UStoreLayerDepth_%d" % k for k in range(orc.maxLayers)]
    for t in range(kw["steps"]):
        P, PET, T = forc.step(t)
        orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
        pre = {k: orc[k].copy() for k in pre_names}
        orc.step()
        out = ref.step(pre, {k: orc[k] for k in gen_golden.PHASE_DIAGS})
        for f in outs:
            assert np.array_equal(orc[f][mask], out[f][mask]), (t, f)
