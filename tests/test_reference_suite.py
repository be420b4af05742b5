"""CPU: the reference's OWN test-suite (qcat/test/test_barcode.py, 14 tests) passes on the replacement.

qcat_b200.dropin grafts GpuScannerMixin onto the reference's scanner classes exactly as on a GPU box; only the device
plan behind it is replaced by the CPU oracle (tests only), so this checks everything between qcat's API and the C ABI
-- kit selection, window packing, layout subsets, record -> dict conversion with the reference's own objects -- against
the assertions the reference's authors wrote.  Needs /root/reference (the fixtures are not shipped anywhere else)."""
import importlib
import os
import sys

import pytest

from tests import helpers

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

HAVE_TESTS = os.path.isfile(os.path.join(refloader.REFERENCE_ROOT, "qcat", "test", "test_barcode.py"))
pytestmark = pytest.mark.skipif(not HAVE_TESTS, reason="reference checkout with its test-suite not available")


OraclePlan = helpers.OraclePlan


def _reference_tests():
    if not HAVE_TESTS:
        return []
    refloader.load()
    module = importlib.import_module("qcat.test.test_barcode")
    return sorted(name for name in dir(module) if name.startswith("test_") and callable(getattr(module, name)))


@pytest.fixture()
def dropin_over_oracle(monkeypatch):
    refloader.load()
    from qcat_b200 import dropin
    from qcat_b200.scanner import GpuScannerMixin, _config_key
    from qcat_b200.tables import Tables
    plans = {}

    def plan_for(self, qcat_config, layouts=None):
        layouts = self.layouts if layouts is None else layouts
        key = (tuple(id(l) for l in layouts), _config_key(qcat_config), float(self.min_quality), self._mode_name())
        if key not in plans:
            plans[key] = OraclePlan(Tables(layouts, qcat_config, self._mode_name(), self.min_quality, getattr(self, "barcodes", None)))
        return plans[key]

    monkeypatch.setattr(GpuScannerMixin, "_plan_for", plan_for)
    dropin.uninstall()
    dropin.install()
    yield plans
    dropin.uninstall()


@pytest.mark.parametrize("name", _reference_tests())
def test_reference_test_passes_on_the_dropin(dropin_over_oracle, name, monkeypatch):
    module = importlib.import_module("qcat.test.test_barcode")
    monkeypatch.chdir(refloader.REFERENCE_ROOT)                 # fixture paths are relative to the checkout root
    getattr(module, name)()
    uses_scanner = name.startswith(("test_barcode", "test_scanner_detect", "test_full_run", "test_trimming"))
    if uses_scanner:
        assert sum(p.calls for p in dropin_over_oracle.values()) > 0, "the replacement was not exercised"
