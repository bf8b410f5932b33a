"""PCRaster CSF version-2 ``.map`` reader / writer (pure numpy).

The reference reads and writes every static map, state and output through PCRaster
(``pcr.readmap`` / ``pcr.report``; ``wf_DynamicFramework.py:1780-1827, 2065-2129, 3261-3520``).  The
format is a 256-byte header followed by the row-major, north-up cell data:

    0   char[32]  "RUU CROSS SYSTEM MAP FORMAT"
    32  u16 version (2), 34 u32 gisFileId, 38 u16 projection (0 y increases down, 1 up),
    40  u32 attrTable, 44 u16 mapType (1), 46 u32 byteOrder (1)
    64  u16 valueScale, 66 u16 cellRepr, 68 8-byte min, 76 8-byte max,
    84  f64 xUL, 92 f64 yUL, 100 u32 nrRows, 104 u32 nrCols, 108 f64 cellSizeX, 116 f64 cellSizeY,
    124 f64 angle
"""
import struct

import numpy as np

SIGNATURE = b"RUU CROSS SYSTEM MAP FORMAT"

VS_BOOLEAN, VS_NOMINAL, VS_ORDINAL, VS_SCALAR, VS_DIRECTION, VS_LDD = 0xE0, 0xE2, 0xF2, 0xEB, 0xFB, 0xF0
CR_UINT1, CR_INT4, CR_REAL4, CR_REAL8, CR_INT1, CR_INT2, CR_UINT2, CR_UINT4 = 0x00, 0x26, 0x5A, 0xDB, 0x04, 0x15, 0x11, 0x22

_DTYPES = {CR_UINT1: np.uint8, CR_INT4: np.int32, CR_REAL4: np.float32, CR_REAL8: np.float64, CR_INT1: np.int8,
           CR_INT2: np.int16, CR_UINT2: np.uint16, CR_UINT4: np.uint32}
_MV = {CR_UINT1: 255, CR_INT4: -2147483648, CR_INT1: -128, CR_INT2: -32768, CR_UINT2: 65535, CR_UINT4: 4294967295}


class MapInfo:
    def __init__(self, nrow, ncol, x_ul=0.0, y_ul=0.0, cellsize=1.0, projection=1, value_scale=VS_SCALAR,
                 cell_repr=CR_REAL4):
        self.nrow, self.ncol, self.x_ul, self.y_ul = int(nrow), int(ncol), float(x_ul), float(y_ul)
        self.cellsize, self.projection = float(cellsize), int(projection)
        self.value_scale, self.cell_repr = value_scale, cell_repr

    def same_grid(self, other):
        return (self.nrow, self.ncol) == (other.nrow, other.ncol)

    def ycoordinate(self):
        """Cell-centre y coordinate per row (pcr.ycoordinate), north-up."""
        sign = -1.0 if self.projection == 1 else 1.0
        return self.y_ul + sign * (np.arange(self.nrow) + 0.5) * self.cellsize


def read_map(path, mv=None):
    """Returns (array, MapInfo).  Scalar maps come back as float32 with NaN for missing values
    (or ``mv`` if given); integer-valued maps as int32 with ``mv`` (default -999 like pcr2numpy)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if buf[:27] != SIGNATURE:
        raise ValueError("%s is not a PCRaster CSF map" % path)
    projection = struct.unpack_from("<H", buf, 38)[0]
    value_scale, cell_repr = struct.unpack_from("<HH", buf, 64)
    x_ul, y_ul = struct.unpack_from("<dd", buf, 84)
    nrow, ncol = struct.unpack_from("<II", buf, 100)
    cellsize = struct.unpack_from("<d", buf, 108)[0]
    if cell_repr not in _DTYPES:
        raise ValueError("%s: unsupported cell representation 0x%x" % (path, cell_repr))
    dt = np.dtype(_DTYPES[cell_repr]).newbyteorder("<")
    data = np.frombuffer(buf, dtype=dt, count=nrow * ncol, offset=256).reshape(nrow, ncol)
    info = MapInfo(nrow, ncol, x_ul, y_ul, cellsize, projection, value_scale, cell_repr)
    if cell_repr in (CR_REAL4, CR_REAL8):
        out = data.astype(np.float32)  # MV is the all-ones bit pattern = a NaN
        if mv is not None:
            out = np.where(np.isnan(out), np.float32(mv), out).astype(np.float32)
        return out, info
    fill = -999 if mv is None else mv
    out = data.astype(np.int64)
    out[data == _MV[cell_repr]] = fill
    return out.astype(np.int32), info


def write_map(path, array, info, value_scale=None, mv=-999.0):
    """Write a map; float arrays as REAL4 scalar, integer arrays as INT4 (nominal) or UINT1 (ldd /
    boolean).  Values equal to ``mv`` or NaN become missing values."""
    a = np.asarray(array)
    if a.shape != (info.nrow, info.ncol):
        raise ValueError("array shape %s does not match the map grid %dx%d" % (a.shape, info.nrow, info.ncol))
    if value_scale is None:
        value_scale = VS_SCALAR if a.dtype.kind == "f" else VS_NOMINAL
    if value_scale in (VS_SCALAR, VS_DIRECTION):
        cell_repr = CR_REAL4
        d = a.astype("<f4").copy()
        miss = ~np.isfinite(d) | (d == np.float32(mv))
        valid = d[~miss]
        d.view("<u4")[miss] = 0xFFFFFFFF
        mn = struct.pack("<f", valid.min()) + b"\xff" * 4 if valid.size else b"\xff" * 8
        mx = struct.pack("<f", valid.max()) + b"\xff" * 4 if valid.size else b"\xff" * 8
    elif value_scale in (VS_LDD, VS_BOOLEAN):
        cell_repr = CR_UINT1
        miss = (a == mv) | (a < 0) | (a > 254)
        d = np.where(miss, 255, a).astype("<u1")
        valid = d[~miss]
        mn = bytes([int(valid.min())]) + b"\xff" * 7 if valid.size else b"\xff" * 8
        mx = bytes([int(valid.max())]) + b"\xff" * 7 if valid.size else b"\xff" * 8
    else:
        cell_repr = CR_INT4
        miss = a == mv
        d = np.where(miss, _MV[CR_INT4], a).astype("<i4")
        valid = d[~miss]
        mn = struct.pack("<i", int(valid.min())) + b"\x00" * 4 if valid.size else b"\xff" * 8
        mx = struct.pack("<i", int(valid.max())) + b"\x00" * 4 if valid.size else b"\xff" * 8
    hdr = bytearray(256)
    hdr[0:27] = SIGNATURE
    struct.pack_into("<H", hdr, 32, 2)
    struct.pack_into("<I", hdr, 34, 0)
    struct.pack_into("<H", hdr, 38, info.projection)
    struct.pack_into("<I", hdr, 40, 0)
    struct.pack_into("<H", hdr, 44, 1)
    struct.pack_into("<I", hdr, 46, 1)
    struct.pack_into("<HH", hdr, 64, value_scale, cell_repr)
    hdr[68:76] = mn
    hdr[76:84] = mx
    struct.pack_into("<dd", hdr, 84, info.x_ul, info.y_ul)
    struct.pack_into("<II", hdr, 100, info.nrow, info.ncol)
    struct.pack_into("<dd", hdr, 108, info.cellsize, info.cellsize)
    struct.pack_into("<d", hdr, 124, 0.0)
    with open(path, "wb") as fh:
        fh.write(bytes(hdr))
        fh.write(d.tobytes())
