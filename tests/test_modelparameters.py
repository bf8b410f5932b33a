"""[modelparameters] of the ini (wf_updateparameters, reference wf_DynamicFramework.py:708-874, 1624-1690): every
supported type on a synthetic case directory, with the CPU oracle behind the framework (host logic, no GPU)."""
import os

import numpy as np
import pytest


@pytest.fixture()
def case(tmp_path, monkeypatch):
    import wflow_b200.framework as fw
    from tests.oracle_backend import OracleSbmModel
    from wflow_b200 import csf, synth
    monkeypatch.setattr(fw, "SbmModel", OracleSbmModel)
    d = str(tmp_path / "case")
    synth.write_case_dir(d, nsteps=4)
    _, info = csf.read_map(os.path.join(d, "staticmaps", "wflow_dem.map"))
    shape = (info.nrow, info.ncol)
    csf.write_map(os.path.join(d, "staticmaps", "myroot.map"), np.full(shape, 123.0, np.float32), info)
    for step in (1, 2, 3, 4):                               # a map stack read every step
        csf.write_map(fw.generate_name_t(os.path.join(d, "inmaps", "CM"), step), np.full(shape, 0.1 * step, np.float32), info)
    for month in (1, 12):                                   # monthly climatology (the run is in January; the clock
        a = np.full(shape, float(month), np.float32)        # starts one step before starttime: 31 December)
        a[0, 0] = np.nan                                    # covered with the default
        csf.write_map(fw.generate_name_t(os.path.join(d, "inmaps", "LAI"), month), a, info)
    with open(os.path.join(d, "intbl", "Kext.tbl"), "w") as fh:   # lookup on one map: the land use
        fh.write("[1,1] 0.5\n[2,> 0.7\n")
    with open(os.path.join(d, "intbl", "KsatVer2.tbl"), "w") as fh:  # landuse, subcatchment, soil
        fh.write("<,> [1,1] <,> 111\n<,> [2,2] <,> 222\n")
    with open(os.path.join(d, "wflow_sbm.ini"), "a") as fh:
        fh.write("\n[modelparameters]\n"
                 "RootingDepth=staticmaps/myroot.map,staticmap,100,1\n"
                 "KsatVer=intbl/KsatVer2.tbl,statictbl,500,1\n"
                 "Cmax=inmaps/CM,timeseries,1.0,0\n"
                 "LAI=inmaps/LAI,monthlyclim,2.5,1\n"
                 "Kext=intbl/Kext.tbl,tbl,0.6,1,staticmaps/wflow_landuse.map\n"
                 "Sl=intbl/nosuchtable.tbl,tbl,0.04,1,staticmaps/wflow_landuse.map\n"
                 "Swood=intbl/nosuchtable.tbl,tbl,0.25,1,staticmaps/wflow_landuse.map\n"
                 "Bogus=too,short\n")
    return d


def test_every_parameter_type_reaches_the_model(case):
    from wflow_b200 import csf
    from wflow_b200.framework import WflowSbm
    m = WflowSbm(Dir=case, RunDir="run1")
    m.createRunId()
    m._runInitial()
    m._runResume()
    assert [p["name"] for p in m._modelparameters] == ["RootingDepth", "KsatVer", "Cmax", "LAI", "Kext", "Sl", "Swood"]
    inside = m._mod._mask
    get = lambda k: m._mod.get(k)
    assert np.all(get("RootingDepth")[inside] == 123.0)                  # staticmap
    sub, _ = csf.read_map(os.path.join(case, "staticmaps", "wflow_subcatch.map"))
    # statictbl replaces the raster BEFORE the derivations: KsatVer is scaled by timestepsecs / basetimestep (= 1 here)
    assert np.all(get("KsatVer")[inside & (sub == 1)] == 111.0) and np.all(get("KsatVer")[inside & (sub == 2)] == 222.0)
    lu, _ = csf.read_map(os.path.join(case, "staticmaps", "wflow_landuse.map"))
    assert np.allclose(get("Kext")[inside & (lu == 1)], 0.5) and np.allclose(get("Kext")[inside & (lu >= 2)], 0.7)  # tbl
    assert np.allclose(get("Sl")[inside], 0.04) and np.allclose(get("Swood")[inside], 0.25)  # missing table: default
    # the clock of dynamic() starts one step before starttime (1 January): 31 December for the first step
    assert m._climate_datetime(1).month == 12 and m._climate_datetime(2).month == 1
    for step, month in ((1, 12.0), (2, 1.0), (3, 1.0), (4, 1.0)):
        m._runDynamic(step, step)
        assert np.allclose(get("Cmax")[inside], 0.1 * step)             # timeseries: a new map every step
        lai = get("LAI")
        assert np.all(lai[inside][1:] == month)                          # monthlyclim
        if inside[0, 0]:
            assert lai[0, 0] == 2.5                                      # missing value -> default
    m._runSuspend()
    m._wf_shutdown()


def test_unsupported_time_dependent_parameter_is_refused_loudly(case, caplog):
    from wflow_b200.framework import WflowSbm
    with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:
        fh.write("M=inmaps/M,monthlyclim,300,1\n")   # M feeds derived statics (f, Kh): not updatable per step
    m = WflowSbm(Dir=case, RunDir="run2")
    m.createRunId()
    import logging
    with caplog.at_level(logging.WARNING, logger="wflow_b200"):
        m._runInitial()
    assert any("M: time-dependent values of this parameter are not supported" in r.message for r in caplog.records)
    m._wf_shutdown()


def test_variable_change_sections(case):
    """[variable_change_once] scales a parameter before the derivations, [variable_change_timestep] the forcing of every
    step (wf_multparameters, wf_DynamicFramework.py:622-690)."""
    from wflow_b200.framework import WflowSbm
    ref = WflowSbm(Dir=case, RunDir="run_a")
    ref.createRunId(); ref._runInitial(); ref._runResume()
    with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:
        fh.write("\n[variable_change_once]\nself.PathFrac = self.PathFrac * 2.0\nself.M = self.M * 0.5\nself.Nonsense = self.Nonsense * 3\n"
                 "\n[variable_change_timestep]\nself.Precipitation = self.Precipitation * 3\n")
    m = WflowSbm(Dir=case, RunDir="run_b")
    m.createRunId(); m._runInitial(); m._runResume()
    inside = m._mod._mask
    assert np.allclose(m._mod.get("PathFrac")[inside], 2.0 * ref._mod.get("PathFrac")[inside])
    # f = (thetaS - thetaR) / M is derived from the scaled M (wflow_sbm.py:2159)
    assert np.allclose(m._mod.get("f")[inside], 2.0 * ref._mod.get("f")[inside], rtol=1e-6)
    ref._mod.request("Precipitation"); m._mod.request("Precipitation")
    ref._runDynamic(1, 1); m._runDynamic(1, 1)
    assert np.allclose(m._mod.get("Precipitation")[inside], 3.0 * ref._mod.get("Precipitation")[inside])
    assert ref._mod.get("Precipitation")[inside].max() > 0
    ref._wf_shutdown(); m._wf_shutdown()


def test_output_rows_labelled_by_datetime(case):
    """[outputcsv_N] timeformat = datetime (wf_DynamicFramework.py:493-496)."""
    from wflow_b200.framework import WflowSbm
    import configparser
    cp = configparser.ConfigParser(); cp.optionxform = str
    cp.read(os.path.join(case, "wflow_sbm.ini"))
    n = 0
    while cp.has_section("outputcsv_%d" % n):  # the sections are numbered consecutively (:1722-1778)
        n += 1
    with open(os.path.join(case, "wflow_sbm.ini"), "a") as fh:
        fh.write("\n[outputcsv_%d]\nsamplemap = staticmaps/wflow_subcatch.map\nfunction = total\ntimeformat = datetime\nself.Precipitation = ptot.csv\n" % n)
    m = WflowSbm(Dir=case, RunDir="run_c")
    m.createRunId(); m._runInitial(); m._runResume()
    m._runDynamic(1, 2)
    m._runSuspend(); m._wf_shutdown()
    rows = open(os.path.join(case, "run_c", "ptot.csv")).read().splitlines()
    assert rows[0].startswith("# Timestep,") and rows[1].startswith("1999-12-31 00:00:00,") and rows[2].startswith("2000-01-01 00:00:00,")


def test_command_line_like_the_reference(case):
    """wflow_sbm.py:3559-3645: -S / -T are times, -P / -p feed the variable_change sections, -I forces a cold start."""
    from wflow_b200 import framework
    rc = framework.main(["-C", case, "-R", "cli", "-S", "2000-01-02 00:00:00", "-T", "2000-01-03 00:00:00", "-I",
                         "-P", "self.PathFrac=self.PathFrac * 2.0", "-p", "self.Precipitation=self.Precipitation * 0.5", "-f"])
    assert rc == 0
    import configparser
    cp = configparser.ConfigParser(); cp.optionxform = str
    cp.read(os.path.join(case, "cli", "runinfo", "configofrun.ini"))
    assert cp.get("run", "starttime") == "2000-01-02 00:00:00" and cp.get("run", "reinit") == "1"
    assert cp.get("variable_change_once", "self.PathFrac") == "self.PathFrac * 2.0"
    rows = [r for r in open(os.path.join(case, "cli", "specrun.csv")).read().splitlines() if not r.startswith("#")]
    assert len(rows) == 2  # two timesteps between the two times
