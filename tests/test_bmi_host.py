"""BMI mirror, host side (no GPU: the CPU oracle behind the framework): attributes, run period, the update loop of the
reference's tests/test_BMI.py (testbmifuncs :136-152, testbmirunnetcdf :203-229)."""
import os

import numpy as np
import pytest


@pytest.fixture()
def ini(tmp_path, monkeypatch):
    import wflow_b200.framework as fw
    from tests.oracle_backend import OracleSbmModel
    from wflow_b200 import synth
    monkeypatch.setattr(fw, "SbmModel", OracleSbmModel)
    d = str(tmp_path / "case")
    synth.write_case_dir(d, nsteps=6)
    return os.path.join(d, "wflow_sbm.ini")


def test_attributes_are_the_ini_options(ini):
    from wflow_b200.bmi import wflowbmi_csdms
    b = wflowbmi_csdms()
    b.initialize_config(ini)
    names = b.get_attribute_names()
    assert "run:starttime" in names and "model:ModelSnow" in names
    b.set_attribute_value(names[0], "SET By TEST")
    assert b.get_attribute_value(names[0]) == "SET By TEST"
    with pytest.raises(Warning):
        b.get_attribute_value("no-colon")
    with pytest.raises(Warning):
        b.set_attribute_value("a:b:c", "x")


def test_set_start_and_end_time_then_run_to_the_end(ini):
    from wflow_b200.bmi import wflowbmi_csdms
    b = wflowbmi_csdms()
    b.initialize_config(ini)
    st0, ts = b.get_start_time(), b.get_time_step()
    b.set_start_time(st0 + ts)            # one day later
    b.set_end_time(st0 + 4 * ts)
    st, ett = b.get_start_time(), b.get_end_time()
    assert st == st0 + ts and ett == st0 + 4 * ts and b.dynModel.nrTimeSteps == 3
    b.initialize_model()
    curtime, cnt = st, 0
    while curtime < ett:
        pet = b.get_value("PET")
        b.set_value("PET", np.where(pet > -999, pet + 10.0, pet))
        b.update_until(curtime + ts)
        cnt += 1
        assert b.get_current_time() - curtime == ts
        curtime = b.get_current_time()
    b.finalize()
    assert cnt == 3 and b.get_current_time() == ett
