"""Stub-loader for the reference's numba kernels (TEST INFRASTRUCTURE ONLY).

Loads ``wflow/wflow_funcs.py`` and ``wflow/wflow_sbm.py`` from the read-only
reference checkout *by file path*, with stand-ins for the modules the reference
imports but this container does not have: ``pcraster`` is the float32-numpy shim of
oracle/pcrshim.py (so the reference's map algebra runs too, unmodified), osgeo /
wflow_adapt are empty.  Nothing is copied; the reference files are executed
where they lie.  Only usable where ``/root/reference`` exists (this container) —
never on the GPU box, and never from the product path.

Used by ``oracle/gen_golden.py`` (to write tests/golden/*.npz) and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).
"""
import importlib.util
import os
import sys
import types

# the reference checkout (build container), or the copy of the three files this loader executes that
# __graft_entry__.build() puts under the git-ignored baseline/_ref/ so that the GPU box can time the reference itself
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("WFLOW_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "wflow", "wflow_sbm.py")):
    REFERENCE_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "wflow", "wflow_sbm.py"))


_cache = {}


def load():
    """Return (wflow_funcs, wflow_sbm) reference modules."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # reference mount is read-only

    def _stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    if "pcraster" not in sys.modules:
        # the float32-numpy shim (oracle/pcrshim.py): the reference's map algebra runs unmodified on it
        from . import pcrshim
        sys.modules["pcraster"] = pcrshim
        fw = _stub(
            "pcraster.framework",
            DynamicModel=type("DynamicModel", (), {"__init__": lambda s: None}),
        )
        pcrshim.framework = fw
    for name in ("osgeo", "osgeo.gdal", "osgeo.gdalconst", "osgeo.ogr", "osgeo.osr"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["osgeo"].gdal = sys.modules["osgeo.gdal"]
    if "wflow" not in sys.modules:
        _stub("wflow").__path__ = []
        _stub("wflow.wflow_adapt")

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    # wflow_lib holds sum_list_cover / idtoid, which wflow_sbm receives through
    # "from wflow.wf_DynamicFramework import *" (the framework itself is not loaded: it needs osgeo, netCDF4 ...)
    wlib = _load("wflow.wflow_lib", os.path.join(REFERENCE_ROOT, "wflow", "wflow_lib.py"))
    if "wflow.wf_DynamicFramework" not in sys.modules:
        _stub("wflow.wf_DynamicFramework", **{k: v for k, v in vars(wlib).items() if not k.startswith("_")})
    funcs = _load("wflow.wflow_funcs", os.path.join(REFERENCE_ROOT, "wflow", "wflow_funcs.py"))
    sbm = _load("wflow.wflow_sbm", os.path.join(REFERENCE_ROOT, "wflow", "wflow_sbm.py"))
    _cache["mods"] = (funcs, sbm)
    return funcs, sbm


def load_hbv():
    """The reference's wflow_hbv module (wflow/wflow_hbv.py), executed where it lies like wflow_sbm above."""
    if "hbv" in _cache:
        return _cache["hbv"]
    load()
    path = os.path.join(REFERENCE_ROOT, "wflow", "wflow_hbv.py")
    if not os.path.isfile(path):
        raise RuntimeError("wflow_hbv.py not present under %s" % REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location("wflow.wflow_hbv", path)
    m = importlib.util.module_from_spec(spec)
    sys.modules["wflow.wflow_hbv"] = m
    spec.loader.exec_module(m)
    _cache["hbv"] = m
    return m
