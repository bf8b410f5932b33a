"""The C-ABI library loads and exports every symbol include/wflow_sbm.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from wflow_b200 import _lib, core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "wflow_sbm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wf_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(l, n), "missing symbol " + n
    assert sorted(_lib.SYMBOLS) == names


def test_field_table_is_consistent():
    l = _lib.lib()
    names = [l.wf_field_name(i).decode() for i in range(l.wf_num_fields())]
    assert len(names) == len(set(names))
    for must in ("P", "PET", "TEMP", "zi", "SubsurfaceFlow", "RiverRunoff", "UStoreLayerDepth_3", "KsatVer", "InwaterMM"):
        assert must in names
    kinds = {n: l.wf_field_kind(i) for i, n in enumerate(names)}
    assert kinds["P"] & 0xff == 0 and kinds["KsatVer"] & 0xff == 1 and kinds["zi"] & 0xff == 2
    assert kinds["RiverRunoff"] & 0x100 and not kinds["LandRunoff"] & 0x100


def test_no_device_fails_loudly():
    """Without a GPU the product must refuse to run rather than fall back to the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(_lib.WflowLibError):
        core.SbmModel(np.full((4, 4), 5, np.uint8))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "wflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "wfo_" not in txt, f
