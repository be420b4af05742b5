"""GPU: the CUDA path against the oracle on RANDOM kit geometries (tests/helpers.random_kit; the same kits
tests/test_custom_layouts.py pins to the reference's Python): adapters from 12 to ~150 nt (beyond the packed adapter
kernel's 104 columns), barcodes of 12-30 nt (beyond the packed barcode kernel's 24 core columns), 1-40 barcodes per
set, shared barcode prefixes / suffixes, sibling layouts -- packed kernels where the plan picks them, generic ones
where it does not, and both forced generic."""
import random

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

_PATHS = {}          # (mode, seed) -> (fast_adapter, fast_barcode) of the plan, checked by the last test


@pytest.fixture(scope="module")
def engine():
    from qcat_b200 import engine as eng
    return eng


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("mode", ["epi2me", "dual"])
def test_cuda_matches_oracle_on_random_kits(engine, mode, seed):
    from qcat_b200 import adapters, config, layout, scanner, synth
    from qcat_b200.tables import Tables
    double = mode == "dual"
    sc = scanner.factory(mode=mode, kit=None if double else "RBK004")
    layouts = helpers.random_kit(random.Random(1000 + seed), layout.AdapterLayout, adapters.Barcode, double)
    data = synth.generate(layouts, 2500, seed=77 + seed, mean_len=500.0, min_len=40, sub=0.05, dele=0.03, ins=0.03)
    tables = Tables(layouts, config.qcatConfig(), mode, sc.min_quality)
    want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
    plan = engine.DevicePlan(tables, device=0)
    info = plan.info()
    _PATHS[(mode, seed)] = (info["fast_adapter"], info["fast_barcode"])
    try:
        for force_generic in (False, True):
            plan.set_force_generic(force_generic)
            got = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
            helpers.assert_records_equal(got, want, "random kit %s/%d generic=%s" % (mode, seed, force_generic))
        # the same reads through the auto-kit entry point with every layout its own kit (per-batch election + restriction)
        kit_names, kit_of_layout = tables.kit_index()
        plan.set_force_generic(False)
        got, batch_kit = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 400,
                                          return_kits=True)
        want_auto, want_kit = helpers.oracle_detect_auto(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"],
                                                         batch_size=400, return_kits=True)
        helpers.assert_records_equal(got, want_auto, "random kit %s/%d auto" % (mode, seed))
        assert np.array_equal(batch_kit, want_kit)
    finally:
        plan.close()
    assert (want["barcode"] >= 0).sum() > 100


def test_random_kits_cover_packed_and_generic_plans():
    """The fuzz is only worth its name if both kernel families were picked by some plan."""
    if len(_PATHS) < 20:
        pytest.skip("the random-kit cases did not all run in this process (test selection / parallel workers)")
    assert {b for _, b in _PATHS.values()} == {0, 1}, _PATHS
    assert 1 in {a for a, _ in _PATHS.values()}, _PATHS
