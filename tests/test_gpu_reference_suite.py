"""GPU: the reference's OWN test-suite (qcat/test/test_barcode.py, 14 tests) passes on the replacement with the real
CUDA plan behind it.

Same as tests/test_reference_suite.py, but nothing is substituted: qcat_b200.dropin grafts GpuScannerMixin onto the
unmodified reference classes (baseline/_ref, shipped with its qcat/test folder by __graft_entry__.build()) and every
detect_barcode / detect_barcode_batch / scan call of the reference's tests runs through libqcat_b200.so on cuda:0."""
import importlib
import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

HAVE_TESTS = os.path.isfile(os.path.join(refloader.REFERENCE_ROOT, "qcat", "test", "test_barcode.py"))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not HAVE_TESTS, reason="reference package with its test-suite not available")]


def _reference_tests():
    if not HAVE_TESTS:
        return []
    refloader.load()
    module = importlib.import_module("qcat.test.test_barcode")
    return sorted(name for name in dir(module) if name.startswith("test_") and callable(getattr(module, name)))


@pytest.fixture()
def dropin_on_gpu():
    refloader.load()
    from qcat_b200 import dropin, engine
    launches = []
    created = []
    original = engine.DevicePlan.__init__

    def tracking_init(self, *args, **kwargs):
        original(self, *args, **kwargs)
        created.append(self)

    engine.DevicePlan.__init__ = tracking_init
    dropin.uninstall()
    dropin.install(device=0)
    try:
        yield created, launches
    finally:
        dropin.uninstall()
        engine.DevicePlan.__init__ = original


@pytest.mark.parametrize("name", _reference_tests())
def test_reference_test_passes_on_the_cuda_plan(dropin_on_gpu, name, monkeypatch):
    created, _ = dropin_on_gpu
    module = importlib.import_module("qcat.test.test_barcode")
    monkeypatch.chdir(refloader.REFERENCE_ROOT)                 # fixture paths are relative to the package root
    getattr(module, name)()
    uses_scanner = name.startswith(("test_barcode", "test_scanner_detect", "test_full_run", "test_trimming"))
    if uses_scanner:
        launched = sum(p.info()["kernel_launches"] for p in created if getattr(p, "_handle", None))
        assert created and launched > 0, "the CUDA plan was not exercised"
