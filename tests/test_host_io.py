"""Host-side I/O of the framework mirror (no GPU): PCRaster CSF maps, lookup tables, map-stack names."""
import os

import numpy as np
import pytest

from wflow_b200 import csf, tbl
from wflow_b200.framework import generate_name_t, lattometres

RHINE = "/root/reference/examples/wflow_rhine_sbm"


def test_csf_roundtrip_scalar_ldd_nominal(tmp_path):
    info = csf.MapInfo(5, 7, x_ul=10.0, y_ul=50.0, cellsize=0.5)
    a = np.arange(35, dtype=np.float32).reshape(5, 7) * 1.5
    a[1, 2] = np.nan
    csf.write_map(tmp_path / "s.map", a, info)
    b, i2 = csf.read_map(tmp_path / "s.map")
    assert i2.same_grid(info) and i2.cellsize == 0.5 and i2.x_ul == 10.0
    assert np.isnan(b[1, 2]) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    ldd = np.full((5, 7), 5, np.int32); ldd[0, 0] = -999; ldd[2, 3] = 7
    csf.write_map(tmp_path / "l.map", ldd, info, value_scale=csf.VS_LDD, mv=-999)
    c, i3 = csf.read_map(tmp_path / "l.map")
    assert i3.value_scale == csf.VS_LDD and c[0, 0] == -999 and c[2, 3] == 7 and c[4, 4] == 5
    ids = np.arange(35, dtype=np.int32).reshape(5, 7) - 3
    ids[0, 0] = -999
    csf.write_map(tmp_path / "n.map", ids, info, value_scale=csf.VS_NOMINAL, mv=-999)
    d, _ = csf.read_map(tmp_path / "n.map")
    assert np.array_equal(d, ids)


@pytest.mark.skipif(not os.path.isdir(RHINE), reason="reference examples not present")
def test_csf_reads_the_reference_rhine_maps():
    ldd, info = csf.read_map(RHINE + "/staticmaps/wflow_ldd.map")
    assert (info.nrow, info.ncol) == (169, 187) and info.value_scale == csf.VS_LDD
    sub, _ = csf.read_map(RHINE + "/staticmaps/wflow_subcatch.map")
    assert (sub > 0).sum() == 15754 and sorted(np.unique(sub[sub > 0])) == list(range(1, 15))
    assert (ldd == 5).sum() >= 1 and set(np.unique(ldd[sub > 0])) <= set(range(1, 10))
    sat, _ = csf.read_map(RHINE + "/instate/SatWaterDepth.map")
    assert np.isfinite(sat[sub > 0]).all() and sat.dtype == np.float32
    # the network of the shipped case: 266 levels
    from wflow_b200 import core
    l = np.where((sub > 0) & (ldd >= 1) & (ldd <= 9), ldd, 0).astype(np.uint8)
    nodes, up, off = core.build_schedule(l)
    assert nodes.size == 15754 and off.size - 1 == 266


def test_tbl_lookup_interval_syntax(tmp_path):
    p = tmp_path / "K.tbl"
    p.write_text("<,7]\t1  1  8000\n<,7] 2 1 100\n[8,> <,> <,> 55\n# comment\n<,> <,> [3,5> 7\n")
    lu = np.array([[1, 7, 8], [9, 3, 3]])
    sc = np.array([[1, 2, 1], [2, 5, 5]])
    so = np.array([[1, 1, 9], [9, 3, 5]])
    out = tbl.lookupscalar(str(p), lu, sc, so)
    assert out[0, 0] == 8000 and out[0, 1] == 100 and out[0, 2] == 55 and out[1, 0] == 55
    assert out[1, 1] == 7 and np.isnan(out[1, 2])


@pytest.mark.skipif(not os.path.isdir(RHINE), reason="reference examples not present")
def test_tbl_reads_the_reference_tables():
    lu, _ = csf.read_map(RHINE + "/staticmaps/wflow_landuse.map")
    sub, _ = csf.read_map(RHINE + "/staticmaps/wflow_subcatch.map")
    soil, _ = csf.read_map(RHINE + "/staticmaps/wflow_soil.map")
    for name in ("KsatVer", "M", "thetaS", "RootingDepth", "CanopyGapFraction", "N"):
        a = tbl.read_tbl_default(RHINE, "intbl", name, lu, sub, soil, -1.0)
        assert np.isfinite(a[sub > 0]).all(), name


def test_map_stack_names_and_cell_lengths():
    assert generate_name_t("inmaps/P", 1).endswith("P0000000.001")
    assert generate_name_t("inmaps/PET", 12).endswith("PET00000.012")
    assert generate_name_t("inmaps/TEMP", 1000).endswith("TEMP0001.000")
    latlen, longlen = lattometres(np.array([0.0, 52.0], np.float32))
    assert abs(latlen[0] - 110574.3) < 1.0 and abs(longlen[1] - 68678.0) < 50.0


def test_area_majority_takes_the_highest_of_tied_values():
    """pcr.areamajority as used by wf_OutputTimeSeriesArea (wf_DynamicFramework.py:477-489)."""
    from wflow_b200.framework import area_majority
    a = np.array([[1, 1, 2, 2], [3, 3, 3, 5], [7, 8, 8, 7]], np.float32)
    ids = np.array([[1, 1, 1, 1], [2, 2, 2, 2], [4, 4, 4, 4]])
    assert list(area_majority(a, ids, [1, 2, 4])) == [2.0, 3.0, 8.0]


def test_run_length_steps_and_intervals():
    """[run] runlengthdetermination (wf_DynamicFramework.py:74-96): 'steps' puts the state time one step before
    starttime (the Rhine example: 29 steps), 'intervals' at starttime itself (one step fewer, first forcing a step later)."""
    import configparser
    import datetime as dt
    from wflow_b200.framework import WflowSbm

    class Holder:
        pass

    for mode, steps, first in (("steps", 29, dt.datetime(1995, 1, 31)), ("intervals", 28, dt.datetime(1995, 2, 1))):
        h = Holder()
        h.config = configparser.ConfigParser()
        h.config.optionxform = str
        h.config.read_string("[run]\nstarttime=1995-01-31 00:00:00\nendtime=1995-02-28 00:00:00\ntimestepsecs=86400\n"
                             "runlengthdetermination=%s\n[model]\n" % mode)
        WflowSbm._read_time(h)
        assert (h.nrTimeSteps, h.datetime_start) == (steps, first)
    h.config.set("run", "runlengthdetermination", "weeks")
    with pytest.raises(ValueError):
        WflowSbm._read_time(h)


def test_tss_round_trip(tmp_path):
    from wflow_b200 import fit
    a = np.array([[1.5, 2.5], [3.0, 1e31], [0.0, 7.25]])
    fit.write_tss(str(tmp_path / "x.tss"), a)
    b = fit.read_tss(str(tmp_path / "x.tss"))
    assert b.shape == (3, 3) and np.array_equal(b[:, 0], [1, 2, 3]) and np.isnan(b[1, 2]) and b[2, 2] == 7.25
