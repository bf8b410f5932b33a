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


def test_oracle_matches_reference_dynamic_live():
    """The reference's whole, unmodified dynamic() -- PCRaster half through oracle/pcrshim.py -- against the C oracle on a
    case that is not among the committed goldens (hourly, modified Rutter, snow, sub-steps 2 / 3): bit for bit."""
    from oracle import ref_dynamic
    from tests import common
    from wflow_b200 import synth
    n = 18
    orc, _, ldd, river = common.make_pair(n, n, kind="forest", layer_thickness=(120, 250), dt=3600, itL=2, itR=3, seed=9,
                                          river_area=6, gpu=False)
    ref = ref_dynamic.RefDynamic(ldd, river, {k: v.copy() for k, v in orc.fields.items()}, timestepsecs=3600,
                                 layer_thickness=(120, 250), kinwaveIters=1, kinwaveLandTstep=1800, kinwaveRiverTstep=1200)
    forc = synth.Forcing(n, n, seed=6, timestepsecs=3600, rain_mean=20.0)
    mask = common.network_mask(orc)
    names = common.state_names(orc.maxLayers) + ["ThroughFall", "StemFlow", "Interception", "SnowMelt", "RainFallPlusMelt",
                                                 "PotSoilEvap", "PotTrans", "InwaterMM", "Inwater", "ActEvap", "SurfaceWatbal"]
    orc.want(*[f for f in names if f not in orc.fields])
    for t in range(4):
        P, PET, T = forc.step(t)
        orc.set("P", P); orc.set("PET", PET); orc.set("TEMP", T)
        orc.step()
        ref.step(P, PET, T)
        for f in names:
            assert np.array_equal(orc[f][mask], ref.get(f)[mask]), (t, f)
