"""PCRaster lookup tables (``pcr.lookupscalar``), as used by ``readtblDefault`` /
``readtblLayersDefault`` (reference ``wflow/wf_DynamicFramework.py:935-1173``).

A ``.tbl`` row holds one key column per input map and the value in the last column.  A key is a
single value or an interval ``[a,b]`` / ``<a,b>`` / ``[a,b>`` / ``<a,b]`` with either bound omitted
(``<,7]``, ``[0,>``, ``<,>``).  The first matching row wins; cells matching no row are missing (NaN).
"""
import os

import numpy as np


def _parse_key(tok):
    tok = tok.strip()
    if tok[0] in "[<" and tok[-1] in "]>":
        lo_closed, hi_closed = tok[0] == "[", tok[-1] == "]"
        lo_s, hi_s = tok[1:-1].split(",")
        lo = float(lo_s) if lo_s.strip() else -np.inf
        hi = float(hi_s) if hi_s.strip() else np.inf
        return lo, hi, lo_closed, hi_closed
    v = float(tok)
    return v, v, True, True


def read_tbl(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.split("#")[0].strip()
            if not line:
                continue
            toks = line.replace("\t", " ").split()
            rows.append(([_parse_key(t) for t in toks[:-1]], float(toks[-1])))
    return rows


def lookupscalar(path, *keymaps):
    """float32 map from a table and its key maps (missing value: NaN)."""
    rows = read_tbl(path)
    shape = np.asarray(keymaps[0]).shape
    out = np.full(shape, np.nan, np.float32)
    todo = np.ones(shape, bool)
    keys = [np.asarray(k, dtype=np.float64) for k in keymaps]
    for conds, val in rows:
        if len(conds) != len(keys):
            raise ValueError("%s: row has %d key columns, %d maps given" % (path, len(conds), len(keys)))
        m = todo.copy()
        for (lo, hi, lc, hc), k in zip(conds, keys):
            m &= (k >= lo) if lc else (k > lo)
            m &= (k <= hi) if hc else (k < hi)
        out[m] = np.float32(val)
        todo &= ~m
    return out


def read_tbl_default(case_dir, intbl, name, landuse, subcatch, soil, default, layer=None):
    """readtblDefault / readtblLayersDefault: ``staticmaps/<name>.map`` if present, else the table,
    else the default (reference wf_DynamicFramework.py:935-1084)."""
    from . import csf
    mapname = os.path.join(case_dir, "staticmaps", name + ".map")
    tblname = os.path.join(case_dir, intbl, name + ".tbl")
    shape = np.asarray(landuse).shape
    if layer is None and os.path.exists(mapname):
        a, _ = csf.read_map(mapname)
        a = a.astype(np.float32)
        return np.where(np.isnan(a), np.float32(default), a).astype(np.float32)
    if os.path.isfile(tblname):
        if layer is None:
            return lookupscalar(tblname, landuse, subcatch, soil)
        return lookupscalar(tblname, landuse, subcatch, soil, np.full(shape, layer))
    return np.full(shape, default, np.float32)
