"""A float32-numpy stand-in for the subset of PCRaster that wflow_sbm's per-timestep path uses
(TEST INFRASTRUCTURE ONLY).

PCRaster cannot be installed in the build container, so the PCRaster half of the reference
(``WflowModel.dynamic`` and the map-algebra functions of ``wflow_funcs.py``) could not be executed and the
C oracle's restatement of it was unpinned.  With this module registered as ``pcraster`` the UNMODIFIED
reference code runs (oracle/refload.py, oracle/ref_dynamic.py): its own operation order, its own branches,
its own constants -- only the primitives below are ours.  Each primitive follows the PCRaster manual's
definition of the operator; where a detail of PCRaster's implementation cannot be verified here it is
listed under "uncertain" so that a reader knows what the pin rests on.

Semantics implemented
  * Scalar maps are REAL4: every operator returns float32 (numpy float32 arithmetic is IEEE single,
    round-to-nearest, like PCRaster's REAL4 loops); Python numbers that meet a map are converted to REAL4
    first (pcraster's operator overloads convert a Python float / int to a nonspatial REAL4 field).
  * Missing values (MV) propagate through every point operator; ``cover`` fills them; ``ifthen`` makes them;
    ``ifthenelse`` is MV where its condition is MV (or where the chosen branch is).
  * Domain errors give MV: x / 0, ln(x <= 0), sqrt(x < 0), x ** y outside pow's real domain.
  * ``pcr2numpy(map, mv)`` / ``numpy2pcr(type, array, mv)`` as in pcraster.numpy_operations: float32 /
    int32 / uint8 arrays, MV <-> mv.
  * ``upstream``, ``catchmenttotal``, ``accucapacitystate``, ``accuflux`` follow the ldd (codes 7 8 9 / 4 5 6 /
    1 2 3, 5 = pit); area operators work per id of a nominal / ordinal / boolean class map.

Uncertain (cannot be checked without PCRaster; none of them is reached with MV-free inputs inside the
model domain, which is what the golden cases use)
  * transcendental functions (``ln``, ``exp``, ``**``, ``sqrt``) are evaluated in float64 and rounded to
    float32 once; PCRaster may call the float32 libm routines directly (difference: at most 1 ulp of REAL4
    on a small fraction of arguments);
  * MV handling of the neighbourhood / area operators when a cell of the area or an upstream cell holds
    MV: here MV cells are skipped by area operators and make ``upstream`` / ``catchmenttotal`` MV downstream.
"""
import numpy as np

F32 = np.float32

# value-scale tags (pcraster.Scalar ...)
Boolean, Nominal, Ordinal, Scalar, Directional, Ldd = "boolean", "nominal", "ordinal", "scalar", "directional", "ldd"
_DTYPE = {Boolean: np.uint8, Nominal: np.int32, Ordinal: np.int32, Scalar: np.float32, Directional: np.float32, Ldd: np.uint8}

_clone = {"shape": None, "cellsize": 1.0}


def setclone(*args):
    """setclone(nrow, ncol[, cellsize, west, north]) -- a map file name is accepted and ignored (the driver sets
    the shape through ``set_shape``)."""
    if len(args) >= 2 and not isinstance(args[0], str):
        _clone["shape"] = (int(args[0]), int(args[1]))
        if len(args) >= 3:
            _clone["cellsize"] = float(args[2])


def set_shape(shape, cellsize=1.0):
    _clone["shape"] = tuple(int(s) for s in shape)
    _clone["cellsize"] = float(cellsize)


def clone_shape():
    if _clone["shape"] is None:
        raise RuntimeError("pcrshim: no clone set (pcrshim.set_shape)")
    return _clone["shape"]


def setglobaloption(_):
    return None


def celllength():
    return Field(np.asarray(_clone["cellsize"], F32), None, Scalar)


class Field:
    """A PCRaster map (spatial: 2-D array; nonspatial: 0-d) with a missing-value mask."""

    __slots__ = ("a", "m", "vs")
    __array_priority__ = 1000  # numpy scalars on the left hand side defer to our reflected operators

    def __init__(self, a, m=None, vs=Scalar):
        a = np.asarray(a, dtype=_DTYPE[vs])
        if m is None:
            m = np.zeros(a.shape, bool)
        else:
            m = np.broadcast_to(np.asarray(m, bool), a.shape)
        self.a, self.m, self.vs = a, m, vs

    # -- helpers ---------------------------------------------------------------------------------------
    @property
    def spatial(self):
        return self.a.ndim > 0

    def __repr__(self):
        return "Field(%s, %s, mv=%d)" % (self.vs, "x".join(map(str, self.a.shape)) or "nonspatial", int(self.m.sum()))

    def __float__(self):
        if self.a.size != 1:
            raise TypeError("only nonspatial fields convert to float")
        if self.m.any():
            raise ValueError("missing value")
        return float(self.a.reshape(-1)[0])

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        raise ValueError("the truth value of a PCRaster field is ambiguous")

    # -- arithmetic ------------------------------------------------------------------------------------
    def __add__(self, o): return _arith(np.add, self, o)
    def __radd__(self, o): return _arith(np.add, o, self)
    def __sub__(self, o): return _arith(np.subtract, self, o)
    def __rsub__(self, o): return _arith(np.subtract, o, self)
    def __mul__(self, o): return _arith(np.multiply, self, o)
    def __rmul__(self, o): return _arith(np.multiply, o, self)
    def __truediv__(self, o): return _div(self, o)
    def __rtruediv__(self, o): return _div(o, self)
    def __pow__(self, o): return _pow(self, o)
    def __rpow__(self, o): return _pow(o, self)
    def __neg__(self): return Field(-_sc(self).a, self.m, Scalar)
    def __pos__(self): return self
    def __abs__(self): return Field(np.abs(_sc(self).a), self.m, Scalar)
    # -- comparisons -> boolean maps -------------------------------------------------------------------
    def __gt__(self, o): return _cmp(np.greater, self, o)
    def __ge__(self, o): return _cmp(np.greater_equal, self, o)
    def __lt__(self, o): return _cmp(np.less, self, o)
    def __le__(self, o): return _cmp(np.less_equal, self, o)
    def __eq__(self, o): return _cmp(np.equal, self, o)
    def __ne__(self, o): return _cmp(np.not_equal, self, o)
    __hash__ = None
    # -- boolean operators -----------------------------------------------------------------------------
    def __and__(self, o): return pcrand(self, o)
    def __rand__(self, o): return pcrand(o, self)
    def __or__(self, o): return pcror(self, o)
    def __ror__(self, o): return pcror(o, self)
    def __invert__(self): return pcrnot(self)


def _as_field(x, like=None):
    """Python / numpy numbers become nonspatial fields of the value scale of ``like`` (default scalar)."""
    if isinstance(x, Field):
        return x
    if isinstance(x, (bool, np.bool_)):
        return Field(np.asarray(1 if x else 0, np.uint8), None, Boolean)
    if isinstance(x, (int, float, np.integer, np.floating)):
        vs = Scalar
        if like is not None and like.vs in (Nominal, Ordinal, Boolean, Ldd) and float(x) == int(x):
            vs = like.vs
        return Field(np.asarray(x, _DTYPE[vs]), None, vs)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return Field(np.asarray(x, F32), None, Scalar)
    raise TypeError("pcrshim: cannot use %r as a PCRaster field" % type(x))


def _sc(x):
    """scalar (REAL4) view of a field or number"""
    x = _as_field(x)
    if x.vs in (Scalar, Directional):
        return x
    return Field(x.a.astype(F32), x.m, Scalar)


def _arith(op, a, b):
    a, b = _sc(a), _sc(b)
    with np.errstate(all="ignore"):
        return Field(op(a.a, b.a, dtype=F32), a.m | b.m, Scalar)


def _div(a, b):
    a, b = _sc(a), _sc(b)
    bad = (b.a == 0)
    with np.errstate(all="ignore"):
        out = np.divide(a.a, np.where(bad, F32(1), b.a), dtype=F32)
    return Field(out, a.m | b.m | bad, Scalar)


def _pow(a, b):
    a, b = _sc(a), _sc(b)
    x, y = a.a.astype(np.float64), b.a.astype(np.float64)
    x, y = np.broadcast_arrays(x, y)
    bad = ((x < 0) & (y != np.floor(y))) | ((x == 0) & (y <= 0))
    with np.errstate(all="ignore"):
        out = np.power(np.where(bad, 1.0, x), y).astype(F32)
    return Field(out, a.m | b.m | bad, Scalar)


def _cmp(op, a, b):
    if not isinstance(a, Field):
        a = _as_field(a, b)
    if not isinstance(b, Field):
        b = _as_field(b, a)
    if a.vs in (Scalar, Directional) or b.vs in (Scalar, Directional):
        a, b = _sc(a), _sc(b)
    return Field(op(a.a, b.a).astype(np.uint8), a.m | b.m, Boolean)


# ---- conversions -----------------------------------------------------------------------------------
def scalar(x):
    x = _as_field(x)
    return Field(x.a.astype(F32), x.m, Scalar)


def boolean(x):
    x = _as_field(x)
    return Field((x.a != 0).astype(np.uint8), x.m, Boolean)


def nominal(x):
    x = _as_field(x)
    return Field(np.trunc(x.a).astype(np.int32) if x.a.dtype.kind == "f" else x.a.astype(np.int32), x.m, Nominal)


def ordinal(x):
    x = _as_field(x)
    return Field(np.trunc(x.a).astype(np.int32) if x.a.dtype.kind == "f" else x.a.astype(np.int32), x.m, Ordinal)


def ldd(x):
    x = _as_field(x)
    return Field(x.a.astype(np.uint8), x.m, Ldd)


def spatial(x):
    x = _as_field(x)
    if x.spatial:
        return x
    sh = clone_shape()
    return Field(np.broadcast_to(x.a, sh).copy(), np.broadcast_to(x.m, sh).copy(), x.vs)


def defined(x):
    x = spatial(_as_field(x))
    return Field((~x.m).astype(np.uint8), None, Boolean)


def numpy2pcr(vs, array, mv):
    a = np.asarray(array)
    if a.ndim != 2:
        raise ValueError("numpy2pcr needs a 2-D array")
    with np.errstate(invalid="ignore"):
        m = (a == mv) if not (isinstance(mv, float) and np.isnan(mv)) else np.isnan(a)
    if a.dtype.kind == "f":
        m = m | np.isnan(a)
    if vs == Boolean:
        out = (np.where(m, 0, a) != 0).astype(np.uint8)
    else:
        with np.errstate(all="ignore"):
            out = np.where(m, 0, a).astype(_DTYPE[vs])
    return Field(out, m, vs)


def pcr2numpy(x, mv):
    x = spatial(_as_field(x))
    out = x.a.copy()
    if x.m.any():
        with np.errstate(all="ignore"):
            out[x.m] = np.asarray(mv).astype(out.dtype) if not (out.dtype.kind in "iu" and isinstance(mv, float) and np.isnan(mv)) else 0
    return out


# ---- point operators -------------------------------------------------------------------------------
def pcrand(a, b):
    a, b = boolean(a), boolean(b)
    return Field(a.a & b.a, a.m | b.m, Boolean)


def pcror(a, b):
    a, b = boolean(a), boolean(b)
    return Field(a.a | b.a, a.m | b.m, Boolean)


def pcrnot(a):
    a = boolean(a)
    return Field(1 - a.a, a.m, Boolean)


def _common(a, b):
    """operands of a two-branch operator in a common value scale"""
    fa, fb = isinstance(a, Field), isinstance(b, Field)
    if fa and not fb:
        b = _as_field(b, a)
    elif fb and not fa:
        a = _as_field(a, b)
    else:
        a, b = _as_field(a), _as_field(b)
    if a.vs != b.vs:
        a, b = _sc(a), _sc(b)
    return a, b


def ifthenelse(cond, a, b):
    cond = boolean(cond)
    a, b = _common(a, b)
    c = cond.a != 0
    val = np.where(c, a.a, b.a)
    m = cond.m | np.where(c, a.m, b.m)
    return Field(val.astype(_DTYPE[a.vs]), m, a.vs)


def ifthen(cond, a):
    cond = boolean(cond)
    a = _as_field(a)
    c = cond.a != 0
    val, m = np.broadcast_arrays(a.a, cond.m | ~c | a.m)
    return Field(val.copy(), m, a.vs)


def cover(*args):
    first = _as_field(args[0])
    out, m, vs = first.a, first.m, first.vs
    if len(args) == 1:  # pcr.cover(0.0): a nonspatial number (wflow_lib.sum_list_cover)
        return first
    for nxt in args[1:]:
        nxt = _as_field(nxt, first)
        if nxt.vs != vs:
            nxt = Field(nxt.a.astype(_DTYPE[vs]), nxt.m, vs)
        out = np.where(m, nxt.a, out)
        m = m & nxt.m
    return Field(out.astype(_DTYPE[vs]), m, vs)


def _minmax(op, args):
    cur = _as_field(args[0])
    if len(args) == 1:
        return cur
    for nxt in args[1:]:
        cur, nxt = _common(cur, nxt)
        cur = Field(op(cur.a, nxt.a).astype(_DTYPE[cur.vs]), cur.m | nxt.m, cur.vs)
    return cur


def max(*args):  # noqa: A001 (pcraster.max)
    return _minmax(np.maximum, args)


def min(*args):  # noqa: A001
    return _minmax(np.minimum, args)


def abs(x):  # noqa: A001
    x = _sc(x)
    return Field(np.abs(x.a), x.m, Scalar)


def _unary64(fn, x, bad=None):
    x = _sc(x)
    v = x.a.astype(np.float64)
    b = np.zeros(v.shape, bool) if bad is None else bad(v)
    with np.errstate(all="ignore"):
        out = fn(np.where(b, 1.0, v)).astype(F32)
    return Field(out, x.m | b, Scalar)


def ln(x):
    return _unary64(np.log, x, lambda v: v <= 0)


def log10(x):
    return _unary64(np.log10, x, lambda v: v <= 0)


def exp(x):
    return _unary64(np.exp, x)


def sqrt(x):
    return _unary64(np.sqrt, x, lambda v: v < 0)


def sqr(x):
    x = _sc(x)
    return _arith(np.multiply, x, x)


def mapminimum(x):
    x = spatial(_as_field(x))
    ok = ~x.m
    if not ok.any():
        return Field(np.asarray(0, _DTYPE[x.vs]), True, x.vs)
    return Field(np.asarray(x.a[ok].min(), _DTYPE[x.vs]), None, x.vs)


def mapmaximum(x):
    x = spatial(_as_field(x))
    ok = ~x.m
    if not ok.any():
        return Field(np.asarray(0, _DTYPE[x.vs]), True, x.vs)
    return Field(np.asarray(x.a[ok].max(), _DTYPE[x.vs]), None, x.vs)


def maptotal(x):
    x = spatial(_sc(x))
    ok = ~x.m
    return Field(np.asarray(x.a[ok].astype(np.float64).sum(), F32), None, Scalar)


# ---- area operators ----------------------------------------------------------------------------------
def _area(fn, x, ids, fill_dtype=F32):
    x, ids = spatial(_sc(x)), spatial(_as_field(ids))
    out = np.zeros(x.a.shape, fill_dtype)
    m = np.ones(x.a.shape, bool)
    ok_id = ~ids.m
    for v in np.unique(ids.a[ok_id]):
        sel = ok_id & (ids.a == v)
        use = sel & ~x.m
        if use.any():
            out[sel] = fn(x.a[use])
            m[sel] = False
    return Field(out, m, Scalar)


def areatotal(x, ids):
    return _area(lambda v: F32(v.astype(np.float64).sum()), x, ids)


def areaaverage(x, ids):
    return _area(lambda v: F32(v.astype(np.float64).sum() / v.size), x, ids)


def areamaximum(x, ids):
    return _area(lambda v: v.max(), x, ids)


def areaminimum(x, ids):
    return _area(lambda v: v.min(), x, ids)


# ---- ldd operators -----------------------------------------------------------------------------------
_DR = {7: (-1, -1), 8: (-1, 0), 9: (-1, 1), 4: (0, -1), 6: (0, 1), 1: (1, -1), 2: (1, 0), 3: (1, 1)}


def _downstream_index(lddf):
    """flat index of the receiver of every cell (-1: pit, MV or pointing outside the map / at an MV cell)"""
    l = spatial(_as_field(lddf))
    nrow, ncol = l.a.shape
    code = np.where(l.m, 0, l.a).astype(np.int64)
    rr, cc = np.meshgrid(np.arange(nrow), np.arange(ncol), indexing="ij")
    ds = np.full((nrow, ncol), -1, np.int64)
    for k, (dr, dc) in _DR.items():
        sel = code == k
        r2, c2 = rr + dr, cc + dc
        ok = sel & (r2 >= 0) & (r2 < nrow) & (c2 >= 0) & (c2 < ncol)
        r2c, c2c = np.clip(r2, 0, nrow - 1), np.clip(c2, 0, ncol - 1)
        ok = ok & (code[r2c, c2c] != 0)
        ds[ok] = (r2c * ncol + c2c)[ok]
    return ds.ravel(), (code.ravel() != 0)


def _topo_order(ds, active):
    """cells ordered so that every cell comes after all of its upstream cells"""
    n = ds.size
    indeg = np.zeros(n, np.int64)
    np.add.at(indeg, ds[(ds >= 0) & active], 1)
    order = np.empty(int(active.sum()), np.int64)
    cur = np.flatnonzero(active & (indeg == 0))
    pos = 0
    while cur.size:
        order[pos:pos + cur.size] = cur
        pos += cur.size
        d = ds[cur]
        d = d[d >= 0]
        np.subtract.at(indeg, d, 1)
        cand = np.unique(d)
        cur = cand[indeg[cand] == 0]
    if pos != order.size:
        raise ValueError("pcrshim: the ldd has a cycle")
    return order


def upstream(lddf, x):
    """sum of x over the cells that drain directly into each cell"""
    x = spatial(_sc(x))
    ds, active = _downstream_index(lddf)
    v = x.a.ravel().astype(F32)
    mv = x.m.ravel()
    out = np.zeros(v.size, np.float64)
    bad = np.zeros(v.size, bool)
    src = np.flatnonzero(active & (ds >= 0))
    # float32 accumulation in a fixed (row-major source) order
    acc = np.zeros(v.size, F32)
    for s in src:  # small test grids only
        acc[ds[s]] = F32(acc[ds[s]] + v[s])
        bad[ds[s]] |= mv[s]
    del out
    return Field(acc.reshape(x.a.shape), (bad | ~active).reshape(x.a.shape), Scalar)


def catchmenttotal(x, lddf):
    """x accumulated over each cell's catchment (the cell itself and everything upstream)"""
    x = spatial(_sc(x))
    ds, active = _downstream_index(lddf)
    order = _topo_order(ds, active)
    acc = x.a.ravel().astype(F32).copy()
    bad = x.m.ravel().copy()
    for c in order:
        d = ds[c]
        if d >= 0:
            acc[d] = F32(acc[d] + acc[c])
            bad[d] |= bad[c]
    return Field(acc.reshape(x.a.shape), (bad | ~active).reshape(x.a.shape), Scalar)


def accuflux(lddf, x):
    return catchmenttotal(x, lddf)


def _accucapacity(lddf, material, capacity):
    material, capacity = spatial(_sc(material)), spatial(_sc(capacity))
    ds, active = _downstream_index(lddf)
    order = _topo_order(ds, active)
    store = material.a.ravel().astype(F32).copy()
    cap = capacity.a.ravel().astype(F32)
    flux = np.zeros(store.size, F32)
    for c in order:
        fl = np.minimum(store[c], cap[c])
        flux[c] = fl
        store[c] = F32(store[c] - fl)
        d = ds[c]
        if d >= 0:
            store[d] = F32(store[d] + fl)
    m = (material.m.ravel() | capacity.m.ravel() | ~active).reshape(material.a.shape)
    return Field(flux.reshape(material.a.shape), m, Scalar), Field(store.reshape(material.a.shape), m, Scalar)


def accucapacitystate(lddf, material, capacity):
    """material left in each cell after every cell passed on min(what it holds, capacity) downstream"""
    return _accucapacity(lddf, material, capacity)[1]


def accucapacityflux(lddf, material, capacity):
    return _accucapacity(lddf, material, capacity)[0]


def downstream(lddf, x):
    x = spatial(_as_field(x))
    ds, active = _downstream_index(lddf)
    v, mv = x.a.ravel(), x.m.ravel()
    idx = np.where(ds >= 0, ds, np.arange(ds.size))
    return Field(v[idx].reshape(x.a.shape), (mv[idx] | ~active).reshape(x.a.shape), x.vs)


def timeinputscalar(*_a, **_k):
    raise NotImplementedError("pcrshim: timeinputscalar (state updating is outside the hot path)")


def readmap(*_a, **_k):
    raise NotImplementedError("pcrshim: readmap -- maps are handed over as arrays (oracle/ref_dynamic.py)")


def report(*_a, **_k):
    return None
