"""CPU: libqcat_b200.so loads, exports every symbol include/qcat_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


@pytest.fixture(scope="module")
def lib():
    from qcat_b200 import build, _ffi
    build.build()
    return _ffi.load()


def test_header_symbols_are_exported(lib):
    from qcat_b200 import _ffi
    header = open(os.path.join(ROOT, "include", "qcat_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qcb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(lib, name), "%s declared in include/qcat_b200.h but not exported" % name
    assert sorted(_ffi.EXPORTS) == declared


def test_struct_layout_matches_header():
    from qcat_b200 import _ffi
    assert ctypes.sizeof(_ffi.QcbResult) == 32
    assert _ffi.RESULT_DTYPE.itemsize == 32
    assert [f[0] for f in _ffi.QcbResult._fields_] == list(_ffi.RESULT_DTYPE.names)
    assert ctypes.sizeof(_ffi.QcbTables) % 8 == 0


def test_version_and_error_channel(lib):
    assert lib.qcb_version().decode().startswith("qcat_b200")
    assert lib.qcb_plan_create(None, 0) is None
    assert b"NULL" in lib.qcb_last_error()


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device plan creation must fail with a clear message (never compute on the CPU)."""
    from qcat_b200 import _ffi, config, engine, scanner
    from qcat_b200.tables import Tables
    if lib.qcb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    with pytest.raises(_ffi.QcbError, match="no CUDA device"):
        engine.DevicePlan(tables, device=0)
    with pytest.raises(_ffi.QcbError, match="no CUDA device"):
        sc.detect_barcode("ACGT" * 100)
    with pytest.raises(_ffi.QcbError, match="no CUDA device"):
        engine.sg_batch(["ACGT"], ["ACGT"], 1, 1, config.qcatConfig().matrix_barcode)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from qcat_b200 import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.load()


def test_table_validation_errors(lib):
    from qcat_b200 import _ffi, config, scanner
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    st, keep = _ffi.tables_struct(tables)
    st.n_layouts = 0
    assert lib.qcb_plan_create(ctypes.byref(st), 0) is None
    assert b"no layouts" in lib.qcb_last_error()
    st, keep = _ffi.tables_struct(tables)
    st.max_align_length = 0
    assert lib.qcb_plan_create(ctypes.byref(st), 0) is None
    assert b"max_align_length" in lib.qcb_last_error()
