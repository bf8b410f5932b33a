"""The CPU oracle against golden vectors produced by the reference's own numba kernels
(tests/golden/*.npz, generator: oracle/gen_golden.py).  Bit-exact: the oracle runs the same
libm on the same operation order."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common


def test_set_dd_golden():
    g = np.load(common.GOLD + "/set_dd.npz")
    names = sorted({k.split("/")[0] for k in g.files})
    assert "quirk2x2" in names
    for n in names:
        net = O.Network(g[n + "/ldd"])
        assert np.array_equal(net.flat_nodes, g[n + "/nodes"]), n
        assert np.array_equal(net.flat_up, g[n + "/nodes_up"]), n
        assert np.array_equal(net.lev_off, g[n + "/lev_off"]), n


def test_set_dd_quirk_drops_index_zero():
    # reference wflow_funcs.py:88 -- `if np.any(idx_up)` is a value test
    net = O.Network(np.array([[6, 5], [5, 5]], np.uint8))
    assert list(net.flat_nodes) == [1, 2, 3]


def test_kinematic_wave_golden():
    g = np.load(common.GOLD + "/kinwave.npz")
    out = np.array([O.kinematic_wave(*a) for a in zip(g["Qin"], g["Qold"], g["q"], g["alpha"], g["beta"], g["dT"], g["dX"])])
    assert np.array_equal(out, g["out"])
    assert (g["out"] == 0).sum() > 0  # the zero branch is covered


def test_kinematic_wave_ssf_golden():
    g = np.load(common.GOLD + "/kinwave_ssf.npz")
    out = np.array([O.kinematic_wave_ssf(*a) for a in zip(*g["args"])])
    assert np.array_equal(out, g["out"], equal_nan=True)


def test_kinematic_wave_ssf_dry_column_golden():
    """Dry and almost dry columns up to f*D ~ 200: the reference leaves before its Newton loop, or sits at its
    1e-30 floor for all 3000 iterations (wflow_funcs.py:224-277).  Bit for bit."""
    g = np.load(common.GOLD + "/kinwave_ssf_dry.npz")
    out = np.array([O.kinematic_wave_ssf(*a) for a in zip(*g["args"])])
    assert np.array_equal(out, g["out"], equal_nan=True)
    assert (g["out"][:, 0] == 1e-30).sum() > 100 and (g["out"][:, 0] == 0).sum() > 100


@pytest.mark.parametrize("name", common.golden_step_cases())
def test_step_trajectory_golden(name):
    g, kw, orc, _ = common.load_golden_case(name)
    mask = common.network_mask(orc)
    outs = sorted({k.split("/")[2] for k in g.files if k.startswith("out/0/")})
    orc.want(*[f for f in outs if f not in orc.fields])
    for t in range(kw["steps"]):
        for f in ("P", "PET", "TEMP"):
            orc.set(f, g["forcing/%d/%s" % (t, f)])
        orc.step()
        for f in outs:
            if f == "SoilWatbal":  # residual of ~1e3 mm terms: summation-order noise of 1e-12 is expected
                assert np.allclose(orc[f][mask], g["out/%d/%s" % (t, f)][mask], rtol=0, atol=1e-9), (name, t, f)
            else:
                assert np.array_equal(orc[f][mask], g["out/%d/%s" % (t, f)][mask]), (name, t, f)


def test_goldens_exercise_the_interesting_branches():
    """Guard against vacuous goldens: exfiltration, infiltration excess, snow, river flow and
    multi-layer unsaturated storage all occur somewhere in the trajectories."""
    seen = dict(exfilt=False, infexc=False, river=False, layers=False, capflux=False)
    for name in common.golden_step_cases():
        g = np.load(common.GOLD + "/step_%s.npz" % name)
        last = max(int(k.split("/")[1]) for k in g.files if k.startswith("out/"))
        for t in range(last + 1):
            seen["exfilt"] |= bool((g["out/%d/ExfiltWater" % t] > 0).any())
            seen["infexc"] |= bool((g["out/%d/InfiltExcess" % t] > 0).any())
            seen["river"] |= bool((g["out/%d/RiverRunoff" % t] > 0).any())
            seen["capflux"] |= bool((g["out/%d/CapFlux" % t] > 0).any())
            if "out/%d/UStoreLayerDepth_2" % t in g.files:
                seen["layers"] |= bool((g["out/%d/UStoreLayerDepth_2" % t] > 0).any())
    assert all(seen.values()), seen
