"""netCDF forcing input / map output of the framework mirror (wflow_b200/ncio.py; reference wflow/wf_netcdfio.py:219-398,
574-790) on the classic netCDF-3 formats: file-level round trips, and a whole run of the framework (CPU oracle behind it)
that must give the same states from a netCDF forcing file as from the PCRaster map stacks it was made from."""
import datetime
import logging
import os

import numpy as np
import pytest

from wflow_b200 import csf, ncio

LOG = logging.getLogger("test")


def _write_forcing_nc(path, stacks, x, y, start, dt, south_first=True, pad=0, fmt=1):
    """A forcing file as pcr2netcdf would write it: time / lat / lon, one variable per stack."""
    from scipy.io import netcdf_file
    nt = len(next(iter(stacks.values())))
    cs = x[1] - x[0]
    xs = np.concatenate([x[0] - cs * np.arange(pad, 0, -1), x, x[-1] + cs * np.arange(1, pad + 1)])
    yn = np.concatenate([y[0] + cs * np.arange(pad, 0, -1), y, y[-1] - cs * np.arange(1, pad + 1)])  # north first
    ds = netcdf_file(path, "w", version=fmt)
    ds.createDimension("time", None)
    ds.createDimension("lat", len(yn))
    ds.createDimension("lon", len(xs))
    t = ds.createVariable("time", "f8", ("time",))
    t.units = "hours since %s" % start.strftime("%Y-%m-%d %H:%M:%S")
    t[:] = np.arange(nt) * dt / 3600.0
    la, lo = ds.createVariable("lat", "f8", ("lat",)), ds.createVariable("lon", "f8", ("lon",))
    la[:] = yn[::-1] if south_first else yn
    lo[:] = xs
    for name, maps in stacks.items():
        v = ds.createVariable(name, "f4", ("time", "lat", "lon"))
        v._FillValue = np.float32(-9999.0)
        for k, a in enumerate(maps):
            full = np.full((len(yn), len(xs)), -9999.0, np.float32)
            full[pad:pad + a.shape[0], pad:pad + a.shape[1]] = np.where(np.isnan(a), -9999.0, a)
            v[k, :, :] = full[::-1] if south_first else full
    ds.close()


def test_input_flip_window_and_time_offset(tmp_path):
    x, y = 100.0 + (np.arange(7) + 0.5) * 2.0, 60.0 - (np.arange(5) + 0.5) * 2.0
    rng = np.random.default_rng(1)
    maps = [rng.random((5, 7)).astype(np.float32) for _ in range(6)]
    maps[2][1, 3] = np.nan
    start = datetime.datetime(2001, 3, 1)
    p = str(tmp_path / "f.nc")
    _write_forcing_nc(p, {"P": maps}, x, y, start, 86400, south_first=True, pad=2)
    nc = ncio.NetcdfInput(p, LOG, ["P", "TEMP"], x, y)
    assert nc.flip and len(nc.datetimelist) == 6 and nc.datetimelist[1] == start + datetime.timedelta(days=1)
    for k in range(6):
        a, ok = nc.gettimestep(k + 1, tsdatetime=start + datetime.timedelta(days=k), var="P")
        assert ok
        np.testing.assert_array_equal(np.isnan(a), np.isnan(maps[k]))
        np.testing.assert_array_equal(np.nan_to_num(a), np.nan_to_num(maps[k]))
    assert nc.gettimestep(1, var="TEMP") == (None, False)      # not in the file: the caller's default
    # a run that starts three days into the file: the offset is found once and kept (wf_netcdfio.py:733-760)
    nc2 = ncio.NetcdfInput(p, LOG, ["P"], x, y)
    a, _ = nc2.gettimestep(1, tsdatetime=start + datetime.timedelta(days=3), var="P")
    np.testing.assert_array_equal(a, maps[3])
    assert nc2.offset == 3
    a, _ = nc2.gettimestep(2, tsdatetime=start + datetime.timedelta(days=4), var="P")
    np.testing.assert_array_equal(a, maps[4])
    with pytest.raises(ValueError, match="do not match model"):
        ncio.NetcdfInput(p, LOG, ["P"], x + 100.0, y)


def test_netcdf4_files_are_refused_with_the_reason(tmp_path):
    p = str(tmp_path / "hdf.nc")
    with open(p, "wb") as fh:
        fh.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(ValueError, match="nccopy -k classic"):
        ncio.NetcdfInput(p, LOG, ["P"], np.arange(3.0), np.arange(3.0)[::-1])


def test_output_round_trip(tmp_path):
    x, y = (np.arange(4) + 0.5), 10.0 - (np.arange(3) + 0.5)
    p = str(tmp_path / "out.nc")
    out = ncio.NetcdfOutput(p, LOG, datetime.datetime(1999, 12, 31), 3, 3600, x, y, maxbuf=2, metadata={"caseName": "c"}, fmt="NETCDF3_64BIT")
    data = [np.arange(12, dtype=np.float32).reshape(3, 4) + 100 * k for k in range(3)]
    data[1][0, 0] = np.nan
    for k, a in enumerate(data):
        out.savetimestep(k + 1, a, var="run", name="self.RiverRunoff", unit="m3/s")
        out.savetimestep(k + 1, a * 2, var="lev", name="self.WaterLevelR")
    out.finish()
    nc = ncio.NetcdfInput(p, LOG, ["run", "lev"], x, y)
    assert not nc.flip and [d.hour for d in nc.datetimelist] == [0, 1, 2]
    for k, a in enumerate(data):
        got, ok = nc.gettimestep(k + 1, var="run")
        assert ok
        np.testing.assert_array_equal(np.nan_to_num(got, nan=-1), np.nan_to_num(a, nan=-1))
        np.testing.assert_array_equal(np.nan_to_num(nc.gettimestep(k + 1, var="lev")[0], nan=-2), np.nan_to_num(a * 2, nan=-2))


def _run(case, run, extra_ini=""):
    from wflow_b200.framework import WflowSbm
    if extra_ini:
        with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:
            fh.write(extra_ini)
    m = WflowSbm(Dir=case, RunDir=run)
    m.createRunId()
    m._runInitial()
    m._runResume()
    m._runDynamic(1, m.nrTimeSteps)
    states = {k: m._get(k) for k in ("zi", "RiverRunoff", "SubsurfaceFlow", "CanopyStorage")}
    m._runSuspend()
    m._wf_shutdown()
    return states


def test_framework_run_from_a_netcdf_forcing_file(tmp_path, monkeypatch):
    import wflow_b200.framework as fw
    from tests.oracle_backend import OracleSbmModel
    from wflow_b200 import synth
    monkeypatch.setattr(fw, "SbmModel", OracleSbmModel)
    case = str(tmp_path / "case")
    nsteps = 4
    synth.write_case_dir(case, nsteps=nsteps)
    want = _run(case, "run_maps")
    # the same forcing as one netCDF-3 file (rows south to north, two cells of margin around the model grid)
    _, info = csf.read_map(os.path.join(case, "staticmaps", "wflow_dem.map"))
    x = info.x_ul + (np.arange(info.ncol) + 0.5) * info.cellsize
    y = info.ycoordinate()
    stacks = {v: [csf.read_map(fw.generate_name_t(os.path.join(case, "inmaps", v), t + 1))[0] for t in range(nsteps)] for v in ("P", "PET", "TEMP")}
    _write_forcing_nc(os.path.join(case, "forcing.nc"), stacks, x, y, datetime.datetime(2000, 1, 1), 86400, south_first=True, pad=2)
    for v in ("P", "PET", "TEMP"):  # the map stacks are gone: the run can only succeed through the file
        for t in range(nsteps):
            os.remove(fw.generate_name_t(os.path.join(case, "inmaps", v), t + 1))
    got = _run(case, "run_nc", "\n[framework]\nnetcdfinput = forcing.nc\nnetcdfoutput = outmaps.nc\nnetcdfwritebuffer = 2\nnetcdf_format = NETCDF3_64BIT\n")
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    assert float(np.abs(want["RiverRunoff"][want["RiverRunoff"] > -999]).max()) > 0
    # [outputmaps] self.RiverRunoff = run  ->  variable "run" of outmaps.nc, equal to the map stack of the first run
    nc = ncio.NetcdfInput(os.path.join(case, "run_nc", "outmaps.nc"), LOG, ["run"], x, y)
    for t in range(nsteps):
        ref, _ = csf.read_map(fw.generate_name_t(os.path.join(case, "run_maps", "outmaps", "run"), t + 1))
        got_t, ok = nc.gettimestep(t + 1, var="run")
        assert ok
        np.testing.assert_array_equal(np.nan_to_num(got_t, nan=-999.0), np.nan_to_num(ref, nan=-999.0))


def test_states_through_netcdf_round_trip(tmp_path, monkeypatch):
    """[framework] netcdfstatesoutput writes the end states as one netCDF-3 file; a second run that resumes from it through
    netcdfstatesinput (reinit = 0, no instate/ maps) starts from exactly the states a run resumed from the map files has."""
    import shutil

    import wflow_b200.framework as fw
    from tests.oracle_backend import OracleSbmModel
    from wflow_b200 import synth
    monkeypatch.setattr(fw, "SbmModel", OracleSbmModel)
    case = str(tmp_path / "case")
    synth.write_case_dir(case, nsteps=3)
    with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:  # summary maps into a netCDF file as well
        fh.write("\n[summary]\nself.RiverRunoff = endrun.map\n[summary_max]\nself.RiverRunoff = maxrun.map\n")
    end = _run(case, "run_a", "\n[framework]\nnetcdfstatesoutput = states.nc\nnetcdfstaticoutput = summary.nc\n")
    _, info = csf.read_map(os.path.join(case, "staticmaps", "wflow_dem.map"))
    sx, sy = info.x_ul + (np.arange(info.ncol) + 0.5) * info.cellsize, info.ycoordinate()
    summ = ncio.NetcdfInput(os.path.join(case, "run_a", "summary.nc"), LOG, ["endrun", "maxrun"], sx, sy)
    er, ok1 = summ.gettimestep(1, var="endrun")
    mr, ok2 = summ.gettimestep(1, var="maxrun")
    assert ok1 and ok2 and not os.path.isfile(os.path.join(case, "run_a", "outsum", "endrun.map"))
    np.testing.assert_array_equal(np.nan_to_num(er, nan=-999.0), end["RiverRunoff"])
    assert (np.nan_to_num(mr, nan=0.0) >= np.nan_to_num(er, nan=0.0)).all()
    assert os.path.isfile(os.path.join(case, "run_a", "states.nc")) and not os.path.isfile(os.path.join(case, "run_a", "outstate", "SatWaterDepth.map"))
    # the same run with map states, for the comparison
    case_b = str(tmp_path / "case_b")
    synth.write_case_dir(case_b, nsteps=3)
    _run(case_b, "run_b")
    shutil.copytree(os.path.join(case_b, "run_b", "outstate"), os.path.join(case_b, "instate"), dirs_exist_ok=True)
    shutil.copyfile(os.path.join(case, "run_a", "states.nc"), os.path.join(case, "instates.nc"))

    def resumed(c, extra):
        text = open(os.path.join(c, "wflow_sbm.ini")).read().replace("reinit = 1", "reinit = 0")
        if "[framework]" in text:
            text = text[: text.index("[framework]")]
        open(os.path.join(c, "wflow_sbm.ini"), "w").write(text + extra)
        m = fw.WflowSbm(Dir=c, RunDir="run_r")
        m.createRunId()
        m._runInitial()
        m._runResume()
        out = {k: m._get(k) for k in m._state_names()}
        m._wf_shutdown()
        return out

    from_nc = resumed(case, "\n[framework]\nnetcdfstatesinput = instates.nc\n")
    from_maps = resumed(case_b, "")
    assert set(from_nc) == set(from_maps) and len(from_nc) >= 14
    for k in from_maps:
        np.testing.assert_array_equal(from_nc[k], from_maps[k], err_msg=k)
    assert float(np.abs(from_maps["SatWaterDepth"][from_maps["SatWaterDepth"] > -999]).max()) > 0


def test_static_maps_through_netcdf_are_refused(tmp_path, monkeypatch):
    import wflow_b200.framework as fw
    from tests.oracle_backend import OracleSbmModel
    from wflow_b200 import synth
    monkeypatch.setattr(fw, "SbmModel", OracleSbmModel)
    case = str(tmp_path / "case")
    synth.write_case_dir(case, nsteps=2)
    with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:
        fh.write("\n[framework]\nnetcdfstaticinput = staticmaps.nc\n")
    m = fw.WflowSbm(Dir=case, RunDir="r")
    m.createRunId()
    with pytest.raises(ValueError, match="netcdfstaticinput"):
        m._runInitial()


CASE_RHINE_HBV = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "wflow_rhine_hbv")


@pytest.mark.skipif(not os.path.isfile(os.path.join(CASE_RHINE_HBV, "wflow_hbv.ini")),
                    reason="examples/wflow_rhine_hbv is not in baseline/_ref (tests/golden/make_ref_cases.py)")
def test_hbv_example_from_a_netcdf_forcing_file(tmp_path, monkeypatch):
    """The shipped wflow_rhine_hbv example (lat / lon grid, warm start, Courant sub-steps) with its map stacks turned into one
    netCDF-3 file: same discharge at every gauge and step as from the map stacks (the CPU restatement behind the host code)."""
    import shutil

    import wflow_b200.framework as fw
    import wflow_b200.hbv as hbv_mod
    from tests.oracle_backend import OracleHbvModel
    from wflow_b200.framework_hbv import WflowHbv
    monkeypatch.setattr(hbv_mod, "HbvModel", OracleHbvModel)

    def run(case, extra=""):
        ini = os.path.join(case, "wflow_hbv.ini")
        if extra:  # the example's ini names netcdfinput = None: set it
            text = open(ini).read()
            assert "netcdfinput = None" in text
            open(ini, "w").write(text.replace("netcdfinput = None", extra, 1))
        m = WflowHbv(Dir=case, RunDir="run_default", configfile="wflow_hbv.ini")
        m.createRunId()
        m._runInitial()
        m._runResume()
        m._runDynamic(1, m.nrTimeSteps)
        m._wf_shutdown()
        return np.genfromtxt(os.path.join(case, "run_default", "run.csv"), delimiter=",")

    a = str(tmp_path / "maps")
    shutil.copytree(CASE_RHINE_HBV, a)
    want = run(a)
    b = str(tmp_path / "nc")
    shutil.copytree(CASE_RHINE_HBV, b)
    _, info = csf.read_map(os.path.join(b, "staticmaps", "wflow_dem.map"))
    x = info.x_ul + (np.arange(info.ncol) + 0.5) * info.cellsize
    y = info.ycoordinate()
    stacks = {v: [csf.read_map(fw.generate_name_t(os.path.join(b, "inmaps", v), t + 1))[0] for t in range(5)] for v in ("P", "PET", "TEMP")}
    _write_forcing_nc(os.path.join(b, "inmaps", "forcing.nc"), stacks, x, y, datetime.datetime(2000, 1, 1), 86400, south_first=False, fmt=2)
    for v in ("P0000000", "PET00000", "TEMP0000"):
        for t in range(1, 11):
            os.remove(os.path.join(b, "inmaps", "%s.%03d" % (v, t)))
    got = run(b, "netcdfinput = inmaps/forcing.nc")
    assert want.shape == (5, 16)
    np.testing.assert_array_equal(got, want)
    assert float(want[:, 2:].max()) > 1500.0
