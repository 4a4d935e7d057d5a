"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/dtb200.h declares, rejects bad arguments without touching a device, and the
host-side helpers behave like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "dtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtb_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from descriptools_b200 import _lib

    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libdtb200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "python binding and header disagree"
    assert _lib.lib.dtb_abi_version() == 1
    assert _lib.lib.dtb_error_string(-3) == b"workspace too small"


def test_argument_validation_needs_no_device():
    from descriptools_b200 import _lib

    L = _lib.lib
    assert L.dtb_slope_d8(None, 0, 10, 10, 0, 10, 12.5, None, None, None) == -1
    assert L.dtb_slope_d8(1, 0, 10, 10, 0, 11, 12.5, 1, None, None) == -1  # row_end > buf_rows
    assert L.dtb_slope_d8(1, 7, 10, 10, 0, 10, 12.5, 1, None, None) == -1  # bad dtype
    assert L.dtb_flowacc(1, 10, 10, 1, 0, -100, 1, 8, None, None) == -3  # workspace too small
    # 1 tile: counters + 256 nodes x (exitw, link, meta, state) + the flat-sweep state used for cyclic grids
    assert L.dtb_flowacc_workspace_bytes(10, 10) == 256 + 256 * (4 + 4 + 4 + 8) + max(10 * 10 * 8, 4096 * 2)
    assert L.dtb_flowacc_workspace_bytes(1000, 1000) == 256 + 256 * 256 * (4 + 4 + 4 + 8) + 1000 * 1000 * 8
    assert L.dtb_hand_workspace_bytes(100, 70) == 256 + 2 * 2 * 256 * 8 + 2 * 2 * 4096 * 2 + 2 * 2 * 4
    assert L.dtb_hand(None, None, 0, None) == -1
    assert L.dtb_downslope(None, 0, None, 1, 1, 1.0, 1.0, 0, None, None) == -1
    with pytest.raises(_lib.DtbError):
        _lib.check(-1, "x")


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import descriptools_b200.slope as slope
    from descriptools_b200 import _lib

    with pytest.raises(_lib.DtbError):
        slope.sloper(np.zeros((4, 4), np.float32), 12.5)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "descriptools_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liborc" not in txt and "libdt_oracle" not in txt, f


def test_divisor_matches_reference_formula():
    from descriptools_b200.helpers import divisor

    br, bc = divisor(2178, 1534, 2, 3)  # helpers.py:5-17
    assert list(br) == [726, 1452] and list(bc) == [383, 767, 1150]
    br, bc = divisor(10, 10, 0, 0)
    assert br.size == 0 and bc.size == 0 and br.dtype == int


def test_dtype_normalisation_is_exact_or_raises():
    from descriptools_b200 import _convert as cv

    assert cv.dem_to_native(np.ones((2, 2), np.int64)).dtype == np.int16
    assert cv.dem_to_native(np.ones((2, 2), np.float64) * 0.5).dtype == np.float32
    with pytest.raises(TypeError):
        cv.dem_to_native(np.full((2, 2), 0.1, np.float64))  # not representable in f32
    assert cv.dem_to_native(np.full((2, 2), 70000, np.int32)).dtype == np.float32
    assert cv.fdr_to_u8(np.array([[-128, 1]], np.int8))[0, 0] == 128
    assert cv.ints_to_native(np.array([[2.0**33]]), "x").dtype == np.int64
    assert cv.ints_to_native(np.array([[5.0, -100.0]]), "x").dtype == np.int32
    with pytest.raises(TypeError):
        cv.ints_to_native(np.array([[0.5]]), "x")
    assert cv.river_to_i8(np.array([[2, 1, 0]])).tolist() == [[0, 1, 0]]
