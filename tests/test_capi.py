"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ftk_c.h declares, and the
product path fails loudly (no CPU fallback) when no GPU is present.  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ftk_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ftk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from feature_tracker_b200 import _capi
    lib = _capi.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_capi.EXPORTED_SYMBOLS) == names
    assert lib.ftk_abi_version() == 4


def test_default_params_are_the_reference_defaults():
    """optical_flow.h:20-28"""
    import ctypes as C
    from feature_tracker_b200 import _capi
    lib = _capi.load_library()
    p = _capi.KltParams()
    lib.ftk_klt_params_default(C.byref(p))
    assert (p.max_track_points, p.max_iteration, p.max_tolerance_large_step, p.patch_row_half, p.patch_col_half, p.method) == (500, 15, 3, 6, 6, 2)
    assert abs(p.max_converge_step - 4e-2) < 1e-9 and list(p.predict) == [1.0, 0.0, 0.0, 1.0]
    import feature_tracker_b200 as ft
    o = ft.OpticalFlowOptions()
    assert (o.kMaxTrackPointsNumber, o.kMaxIteration, o.kMaxToleranceLargeStep, o.kPatchRowHalfSize, o.kMethod) == (500, 15, 3, 6, ft.OpticalFlowMethod.kFast)
    m = ft.MatcherOptions()
    assert (m.kMaxValidPredictRowDistance, m.kMaxValidPredictColDistance, m.kMaxValidDescriptorDistance) == (40, 40, 0.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    import feature_tracker_b200 as ft
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ft.FtkError):
        ft.Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under feature_tracker_b200/ or include/ may reference it."""
    bad = []
    for base in ("feature_tracker_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    if re.search(r"\bimport oracle|from oracle|ftko_|ftkref_|libftk_oracle|libftk_ref|pyoracle", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_pack_brief_layout():
    import numpy as np
    import feature_tracker_b200 as ft
    bits = np.zeros((2, 256), np.uint8)
    bits[0, 0] = bits[0, 33] = bits[1, 255] = 1
    w = ft.pack_brief(bits)
    assert w.shape == (2, 8) and w[0, 0] == 1 and w[0, 1] == 2 and w[1, 7] == 0x80000000
