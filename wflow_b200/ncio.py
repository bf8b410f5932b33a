"""netCDF forcing input and map output of the framework mirror (wflow/wf_netcdfio.py: `netcdfinput` :574-790,
`netcdfoutput` :219-398, `prepare_nc` :51-216), on the classic netCDF-3 format.

The reference goes through the netCDF4 library, which reads and writes netCDF-4 (HDF5) and the classic formats alike
(`[framework] netcdf_format = NETCDF4 | NETCDF3_CLASSIC | NETCDF3_64BIT`).  That library (and HDF5) is not in this image;
`scipy.io.netcdf_file` is, and it speaks the classic formats.  So: forcing is read from classic / 64-bit-offset files
(`nccopy -k classic` converts a netCDF-4 file), output is written as 64-bit-offset netCDF-3 whatever `netcdf_format` says
(logged), and a netCDF-4 input file is refused with that explanation.  Semantics follow the reference: variables are
named by the basename of the `[inputmapstacks]` entries, `time` is matched against the timestep's date / time with the
reference's offset search, rows are flipped when latitude ascends, the window of the model grid is cut out of a larger
file by coordinates, `_FillValue` / `missing_value` become missing cells, `scale_factor` / `add_offset` are applied.
"""
import bisect
import datetime as _dt
import os
import re

import numpy as np

F32 = np.float32
_UNIT_SECONDS = {"second": 1.0, "seconds": 1.0, "sec": 1.0, "secs": 1.0, "s": 1.0, "minute": 60.0, "minutes": 60.0, "min": 60.0,
                 "hour": 3600.0, "hours": 3600.0, "hr": 3600.0, "h": 3600.0, "day": 86400.0, "days": 86400.0, "d": 86400.0}


def _open(path, mode="r", **kw):
    from scipy.io import netcdf_file
    if mode == "r":
        with open(path, "rb") as fh:
            magic = fh.read(4)
        if magic[:3] != b"CDF":
            kind = "a netCDF-4 / HDF5 file" if magic == b"\x89HDF" else "not a netCDF file"
            raise ValueError("%s: %s.  This implementation reads the classic netCDF-3 formats only (the netCDF4 / HDF5 "
                             "libraries are not available); convert with `nccopy -k classic in.nc out.nc`." % (path, kind))
    return netcdf_file(path, mode, mmap=False, **kw)


def parse_time_units(units):
    """'seconds since 1990-01-01 00:00:00.0 00:00' -> (seconds per unit, epoch)."""
    m = re.match(r"\s*(\w+)\s+since\s+(\d{1,4})-(\d{1,2})-(\d{1,2})(?:[ T](\d{1,2}):(\d{1,2})(?::(\d{1,2}(?:\.\d*)?))?)?", units)
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError("time units not understood: %r" % units)
    y, mo, d = int(m.group(2)), int(m.group(3)), int(m.group(4))
    hh, mm = int(m.group(5) or 0), int(m.group(6) or 0)
    ss = float(m.group(7) or 0.0)
    return _UNIT_SECONDS[m.group(1).lower()], _dt.datetime(y, mo, d, hh, mm) + _dt.timedelta(seconds=ss)


def _attr(v, name, default=None):
    a = getattr(v, name, default)
    return a.decode() if isinstance(a, bytes) else a


class NetcdfInput:
    """netcdfinput (wf_netcdfio.py:574-790): x / y are the cell-centre coordinates of the model grid (north-up)."""

    def __init__(self, path, logger, variables, x, y):
        if not os.path.exists(path):
            raise ValueError(os.path.abspath(path) + " not found!")
        self.logger = logger
        self.ds = _open(path)
        logger.info("Reading input from netCDF file: " + path)
        tv = self.ds.variables["time"]
        per, epoch = parse_time_units(_attr(tv, "units", "Seconds since 1970-01-01 00:00:00"))
        cal = str(_attr(tv, "calendar", "gregorian")).lower()
        if cal not in ("gregorian", "standard", "proleptic_gregorian"):
            raise ValueError("%s: calendar %s is not supported (gregorian only)" % (path, cal))
        self.datetimelist = [epoch + _dt.timedelta(seconds=float(t) * per) for t in np.asarray(tv[:], np.float64)]
        self.offset = 0
        v = self.ds.variables
        self.x = np.asarray((v["x"] if "x" in v else v["lon"])[:], np.float64)
        yy = np.asarray((v["y"] if "y" in v else v["lat"])[:], np.float64)
        self.y = yy if yy.ndim == 1 else yy[:][0]
        self.flip = not (self.y[0] > self.y[-1])  # rows ascend with latitude: flip to north-up
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        acc = np.diff(x).mean() * 0.25 if x.size > 1 else 0.25  # (a quarter of a cell: rounding of the coordinates)
        ys = self.y[::-1] if self.flip else self.y
        (self.latidx,) = np.logical_and(ys + acc >= y.min(), ys <= y.max() + acc).nonzero()
        (self.lonidx,) = np.logical_and(self.x + acc >= x.min(), self.x <= x.max() + acc).nonzero()
        if len(self.lonidx) != len(x):
            raise ValueError("X coordinates in netcdf do not match model (model %g to %g, file %g to %g)"
                             % (x.min(), x.max(), self.x.min(), self.x.max()))
        if len(self.latidx) != len(y):
            raise ValueError("Y coordinates in netcdf do not match model (model %g to %g, file %g to %g)"
                             % (y.min(), y.max(), self.y.min(), self.y.max()))
        self.alldat = {}
        for var in variables:
            if var in v:
                self.alldat[var] = v[var]
            else:
                logger.warning("Variable " + var + " not found in netcdf file: " + path)

    def gettimestep(self, timestep, tsdatetime=None, var="P", shifttime=False):
        """(float32 raster with NaN for missing cells, True), or (None, False) when the file does not hold `var`."""
        ncindex = (timestep if shifttime else timestep - 1) + self.offset
        n = len(self.datetimelist)
        if n < ncindex + 1:
            ncindex = n - 1
        if tsdatetime is not None and tsdatetime != self.datetimelist[ncindex]:
            self.logger.warning("Date/time does not match. Wanted %s got %s" % (tsdatetime, self.datetimelist[ncindex]))
            pos = bisect.bisect_left(self.datetimelist, tsdatetime)
            if pos >= n:
                pos = n - 1
                self.logger.warning("No matching date/time found using last date/time again...")
            self.offset += pos - ncindex
            ncindex = pos
        if var not in self.alldat:
            return None, False
        d = self.alldat[var]
        la0, la1 = int(self.latidx.min()), int(self.latidx.max()) + 1
        lo0, lo1 = int(self.lonidx.min()), int(self.lonidx.max()) + 1
        if self.flip:  # latidx counts from the north of the flipped rows
            nlat = self.y.size
            rows = slice(nlat - la1, nlat - la0)
        else:
            rows = slice(la0, la1)
        raw = np.array(d[ncindex, rows, lo0:lo1] if len(d.dimensions) == 3 else d[ncindex, 0, rows, lo0:lo1])
        miss = None
        for name in ("_FillValue", "missing_value"):
            if hasattr(d, name):
                miss = np.asarray(getattr(d, name)).reshape(-1)[0]
                break
        bad = (raw == miss) if miss is not None else np.zeros(raw.shape, bool)
        a = raw.astype(np.float64)
        if hasattr(d, "scale_factor") or hasattr(d, "add_offset"):
            a = a * float(np.asarray(getattr(d, "scale_factor", 1.0)).reshape(-1)[0]) + float(np.asarray(getattr(d, "add_offset", 0.0)).reshape(-1)[0])
        a = np.where(bad, np.nan, a).astype(F32)
        return (np.flipud(a).copy() if self.flip else a), True

    def close(self):
        self.ds.close()


class NetcdfOutput:
    """netcdfoutput (wf_netcdfio.py:219-398): one file for all [outputmaps], a variable per map, written in buffers of
    `maxbuf` timesteps.  x / y: cell-centre coordinates of the model grid; rows are written south to north like the
    reference's files (`y` ascending ... the reference keeps the PCRaster row order: north first -- so do we)."""

    def __init__(self, path, logger, starttime, timesteps, timestepsecs, x, y, *, maxbuf=50, metadata=None, fmt="NETCDF4",
                 fill_value=1e31):
        self.path, self.logger, self.maxbuf = path, logger, max(int(maxbuf), 1)
        self.fill = F32(fill_value)
        if fmt.upper() not in ("NETCDF3_64BIT", "NETCDF3_64BIT_OFFSET"):
            logger.warning("netcdf_format = %s: written as NETCDF3_64BIT (netCDF-3 is what this implementation writes)" % fmt)
        if os.path.exists(path):
            os.remove(path)
        self.nsteps = int(timesteps)
        ds = _open(path, "w", version=2)
        ds.createDimension("time", None)
        ds.createDimension("lat", len(y))
        ds.createDimension("lon", len(x))
        t = ds.createVariable("time", "f8", ("time",))
        t.units = "seconds since %04d-%02d-%02d %02d:%02d:%02d.0 00:00" % (starttime.year, starttime.month, starttime.day,
                                                                         starttime.hour, starttime.minute, starttime.second)
        t.calendar = "gregorian"
        t.standard_name = "time"
        t.long_name = "time"
        t.axis = "T"
        la = ds.createVariable("lat", "f8", ("lat",))
        la.standard_name, la.long_name, la.units, la.axis = "latitude", "latitude", "degrees_north", "Y"
        lo = ds.createVariable("lon", "f8", ("lon",))
        lo.standard_name, lo.long_name, lo.units, lo.axis = "longitude", "longitude", "degrees_east", "X"
        la[:] = np.asarray(y, np.float64)
        lo[:] = np.asarray(x, np.float64)
        for k, val in (metadata or {}).items():
            setattr(ds, k, str(val))
        ds.Conventions = "CF-1.4"
        self.ds = ds
        self._time = t
        self._dt = float(timestepsecs)
        self._vars, self._buf, self._written = {}, {}, 0

    def savetimestep(self, timestep, data, unit="mm", var="P", name="Precipitation"):
        """data: float32 raster, NaN or the fill value where missing; timestep is 1-based."""
        if var not in self._vars:
            self.logger.debug("Creating variable " + var + " in netcdf file.")
            v = self.ds.createVariable(var, "f4", ("time", "lat", "lon"))
            v.units, v.standard_name, v.long_name = unit, name, name
            v._FillValue = self.fill
            self._vars[var] = v
        a = np.asarray(data, F32)
        a = np.where(np.isnan(a), self.fill, a).astype(F32)
        idx = int(timestep) - 1
        self._vars[var][idx, :, :] = a
        if idx >= self._written:
            for k in range(self._written, idx + 1):
                self._time[k] = k * self._dt
            self._written = idx + 1
        if (idx + 1) % self.maxbuf == 0:
            self.ds.flush()

    def finish(self):
        if self.ds is not None:
            self.ds.close()
            self.ds = None
