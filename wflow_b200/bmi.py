"""BMI surface of the hot path: same method names, argument meaning and error behaviour as the
reference's ``wflowbmi_csdms`` (``wflow/wflow_bmi.py:536-1341``, base class ``wflow/bmi.py:32-393``),
on top of `framework.WflowSbm`.

Semantics kept: grids are float32 (nrow, ncol), flipped south-up,
-999 for missing; a length-1 ``src`` in ``set_value`` means a uniform map; only names whose [API]
role is 0 (input), 2 (state) or 3 (parameter) are served by ``get_value`` (others: logged error,
``None``); ``set_value`` on a role-1 (output-only) name raises ``ValueError``; time is seconds since
the epoch, UTC; ``update_until`` may step back exactly one step (device-side state snapshot).
"""
import calendar
import datetime
import logging
import os

import numpy as np

from .framework import WflowSbm, configget


def _utc(seconds):
    return datetime.datetime.fromtimestamp(seconds, datetime.timezone.utc).replace(tzinfo=None)


class wflowbmi_csdms:
    def __init__(self, log=None):
        self.bmilogger = log or logging.getLogger("wflow_b200.bmi")
        self.dynModel = None
        self.currenttimestep = 0
        self.name = "undefined"
        self.myModel = None

    # ---- control ---------------------------------------------------------------------------------
    def initialize_config(self, filename, loglevel=logging.DEBUG):
        """wflow_bmi.py:584-655: read the ini, create the model, do not run initial() yet."""
        self.datadir = os.path.dirname(os.path.abspath(filename))
        inifile = os.path.basename(filename)
        self.dynModel = WflowSbm(Dir=self.datadir, RunDir="run_default", configfile=inifile)
        # wflow_bmi.py:606-640 picks the model module by [model] modeltype
        modeltype = configget(self.dynModel.config, "model", "modeltype", "sbm").strip()
        if modeltype in ("hbv", "wflow_hbv"):
            from .framework_hbv import WflowHbv
            self.dynModel = WflowHbv(Dir=self.datadir, RunDir="run_default", configfile=inifile)
            self.name = "wflow_hbv"
        elif modeltype in ("sbm", "wflow_sbm"):
            self.name = "wflow_sbm"
        else:
            raise ValueError("modeltype = %s: only wflow_sbm and wflow_hbv are implemented" % modeltype)
        self.dynModel.createRunId()
        roles = self.dynModel.wf_supplyVariableNamesAndRoles()
        self.inputoutputvars = [r[0] for r in roles if r[1] in (0, 2, 3)]
        self.outputonlyvars = [r[0] for r in roles if r[1] == 1]
        self.units = {r[0]: r[2] for r in roles}

    def initialize_model(self):
        """wflow_bmi.py:657-668: setupFramework, _runInitial, _runResume."""
        self.dynModel._runInitial()
        self.dynModel._runResume()
        self.currenttimestep = 0

    def initialize(self, filename, loglevel=logging.DEBUG):
        self.initialize_config(filename, loglevel=loglevel)
        self.initialize_model()

    def update(self):
        """One timestep (wflow_bmi.py:791-802)."""
        self.currenttimestep += 1
        self.dynModel._runDynamic(self.currenttimestep, self.currenttimestep)

    def update_until(self, time):
        """wflow_bmi.py:804-845: forward by whole steps, or back by exactly one step."""
        curtime = self.get_current_time()
        if abs(time - curtime) % self.dynModel.timestepsecs != 0:
            raise ValueError("time step not a multiple of the model time step")
        if curtime > time:
            nsteps = int((curtime - time) / self.dynModel.timestepsecs)
            if nsteps != 1:
                raise ValueError("Time before current time, and not exactly one step back.")
            self.dynModel.wf_QuickResume()
            self.currenttimestep -= 1
            self.dynModel._currentTimeStep -= 1
            return
        nsteps = int((time - curtime) / self.dynModel.timestepsecs)
        for _ in range(nsteps):
            self.update()

    def update_frac(self, time_frac):
        raise NotImplementedError("update_frac")  # the reference raises too (wflow_bmi.py:847)

    def save_state(self, destination_directory):
        self.dynModel.wf_suspend(destination_directory)

    def load_state(self, source_directory):
        self.dynModel.wf_resume(source_directory)

    def finalize(self):
        self.dynModel._runSuspend()
        self.dynModel._wf_shutdown()

    # ---- info ---------------------------------------------------------------------------------------
    def get_component_name(self):
        return self.name

    def get_input_var_names(self):
        return [r[0] for r in self.dynModel.wf_supplyVariableNamesAndRoles() if r[1] in (0, 3)]

    def get_output_var_names(self):
        return [r[0] for r in self.dynModel.wf_supplyVariableNamesAndRoles() if r[1] in (1, 2)]

    def get_var_type(self, long_var_name):
        v = self.get_value(long_var_name)
        return str(v.dtype) if v is not None else "float32"

    def get_var_rank(self, long_var_name):
        v = self.get_value(long_var_name)
        return len(v.shape) if v is not None else 2

    def get_var_size(self, long_var_name):
        v = self.get_value(long_var_name)
        return v.size if v is not None else 1

    def get_var_nbytes(self, long_var_name):
        v = self.get_value(long_var_name)
        return v.nbytes if v is not None else 4

    def get_var_units(self, long_var_name):
        return self.units.get(long_var_name, "-")

    # ---- time (seconds since 1970-01-01 UTC) ---------------------------------------------------------
    def get_start_time(self):
        return float(calendar.timegm(self.dynModel.datetime_start.utctimetuple())) - self.dynModel.timestepsecs

    def get_end_time(self):
        return float(calendar.timegm(self.dynModel.datetime_end.utctimetuple()))

    def get_current_time(self):
        return self.get_start_time() + self.currenttimestep * self.dynModel.timestepsecs

    def get_time_step(self):
        return float(self.dynModel.timestepsecs)

    def _set_run_period(self, start=None, end=None):
        m = self.dynModel
        if start is not None:  # get_start_time() is the state time, one step before the first forcing
            m.datetime_start = _utc(start) + datetime.timedelta(seconds=m.timestepsecs)
        if end is not None:
            m.datetime_end = _utc(end)
        m.nrTimeSteps = max(1, int((m.datetime_end - m.datetime_start).total_seconds() // m.timestepsecs) + 1)

    def set_start_time(self, start_time):
        """wflow_bmi.py:675-701: shift the start of the run (seconds since the epoch, UTC)."""
        self._set_run_period(start=start_time)

    def set_end_time(self, end_time):
        """wflow_bmi.py:703-724."""
        self._set_run_period(end=end_time)

    # ---- attributes = options of the ini, 'section:option' (wflow_bmi.py:726-767) ---------------------
    def get_attribute_names(self):
        cfg = self.dynModel.config
        return [sect + ":" + opt for sect in cfg.sections() for opt in cfg.options(sect)]

    def get_attribute_value(self, attribute_name):
        path = attribute_name.split(":")
        if len(path) != 2:
            raise Warning("attributes should follow the name:option  convention")
        return self.dynModel.config.get(path[0], path[1])

    def set_attribute_value(self, attribute_name, attribute_value):
        path = attribute_name.split(":")
        if len(path) != 2:
            self.bmilogger.warning("Attributes should follow the name:option  convention")
            raise Warning("attributes should follow the name:option  convention")
        self.dynModel.config.set(path[0], path[1], attribute_value)

    def get_time_units(self):
        return "seconds since 1970-01-01 00:00:00.0 00:00"

    # ---- values ---------------------------------------------------------------------------------------
    def get_value(self, long_var_name):
        """wflow_bmi.py:1081-1110."""
        if long_var_name in self.inputoutputvars:
            ret = self.dynModel.wf_supplyMapAsNumpy(long_var_name)
            if ret is None:
                self.bmilogger.error("%s not available" % long_var_name)
            return ret
        self.bmilogger.error("%s not in the list of input/output variables ([API] section)" % long_var_name)
        return None

    def get_value_at_indices(self, long_var_name, inds):
        v = self.get_value(long_var_name)
        return None if v is None else v[inds]

    def set_value(self, long_var_name, src):
        """wflow_bmi.py:1291-1320."""
        if long_var_name in self.outputonlyvars:
            raise ValueError("set_value called for a variable that is output only: " + long_var_name)
        if long_var_name not in self.inputoutputvars:
            raise ValueError(long_var_name + " not in the list of input/output variables ([API] section)")
        src = np.asarray(src)
        if src.size == 1:
            self.dynModel.wf_setValues(long_var_name, float(src.reshape(-1)[0]))
        else:
            self.dynModel.wf_setValuesAsNumpy(long_var_name, src)

    def set_value_at_indices(self, long_var_name, inds, src):
        v = self.get_value(long_var_name)
        if v is None:
            raise ValueError(long_var_name + " cannot be read back")
        v = v.copy()
        v[inds] = src
        self.set_value(long_var_name, v)

    # ---- grid (wflow_bmi.py:1130-1290) ----------------------------------------------------------------
    def get_grid_type(self, long_var_name):
        return "uniform_rectilinear"

    def get_grid_shape(self, long_var_name):
        return self.dynModel.shape

    def get_grid_spacing(self, long_var_name):
        cs = self.dynModel._info.cellsize
        return [cs, cs]

    def get_grid_origin(self, long_var_name):
        """(y, x) of the centre of the lower-left cell, like the reference (south-up grids)."""
        info = self.dynModel._info
        y = info.ycoordinate()
        return [float(y[-1]), info.x_ul + 0.5 * info.cellsize]

    def get_grid_x(self, long_var_name):
        info = self.dynModel._info
        x = info.x_ul + (np.arange(info.ncol) + 0.5) * info.cellsize
        return np.tile(x, (info.nrow, 1)).astype(np.float32)

    def get_grid_y(self, long_var_name):
        info = self.dynModel._info
        return np.flipud(np.tile(info.ycoordinate()[:, None], (1, info.ncol))).astype(np.float32)

    def get_grid_z(self, long_var_name):
        return np.zeros(self.dynModel.shape, np.float32)

    def get_grid_connectivity(self, long_var_name):
        raise NotImplementedError  # as the reference does (wflow_bmi.py:1330-1335)

    def get_grid_offset(self, long_var_name):
        raise NotImplementedError  # wflow_bmi.py:1337-1341
