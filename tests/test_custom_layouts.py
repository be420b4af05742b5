"""CPU: the oracle against the UNMODIFIED reference Python on RANDOM kit geometries (flank lengths 0-45, barcode lengths
12-30, 1-40 barcodes, shared barcode prefixes / suffixes, sibling layouts, double-barcode layouts in dual mode) -- the
shipped kits all look alike (24-nt barcodes, ~11-nt contexts), so this is where geometry-dependent indexing in
extract_barcode_region / the context handling (scanner_base.py:29-141) would show.  tests/test_gpu_custom_layouts.py
then holds the CUDA path to the oracle on the same kits."""
import os
import random
import sys

import pytest

from tests import helpers

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

pytestmark = pytest.mark.skipif(not refloader.available(), reason="reference package not available")


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("mode", ["epi2me", "dual"])
def test_oracle_equals_reference_on_random_kits(mode, seed):
    refloader.load()
    from qcat import adapters as ref_adapters
    from qcat import config as ref_config
    from qcat import layout as ref_layout
    from qcat import scanner as ref_scanner
    from qcat_b200 import adapters, config, layout, scanner, synth
    from qcat_b200.tables import Tables, pack_windows
    double = mode == "dual"
    ref = ref_scanner.factory(mode=mode, kit=None if double else "RBK004")
    mine = scanner.factory(mode=mode, kit=None if double else "RBK004")
    ref.layouts = helpers.random_kit(random.Random(1000 + seed), ref_layout.AdapterLayout, ref_adapters.Barcode, double)
    mine.layouts = helpers.random_kit(random.Random(1000 + seed), layout.AdapterLayout, adapters.Barcode, double)
    assert [l.sequence for l in ref.layouts] == [l.sequence for l in mine.layouts]
    data = synth.generate(mine.layouts, 150, seed=seed, mean_len=500.0, min_len=40, sub=0.05, dele=0.03, ins=0.03)
    reads = synth.windows_to_reads(data)
    tables = Tables(mine.layouts, config.qcatConfig(), mode, mine.min_quality)
    win5, tail3, wlen, read_len, _ = pack_windows(reads, 150)
    got = helpers.oracle_detect(tables, win5, tail3, wlen, read_len)
    cfg = ref_config.qcatConfig()
    for read, g in zip(reads, got):
        w = ref.detect_barcode(read, None, cfg)
        if w["barcode"] is None:
            assert g["barcode"] < 0 and int(g["exit_status"]) == w["exit_status"], read[:80]
        else:
            picked = tables.barcode_object(int(g["layout"]), int(g["barcode"]))
            ident = "{}/{}".format(picked[0].id, picked[1].id) if double else picked.id
            assert ident == w["barcode"].id and mine.layouts[int(g["layout"])].sequence == w["adapter"].sequence, read[:80]
            assert float(g["barcode_score"]) == w["barcode_score"] and int(g["adapter_end"]) == w["adapter_end"], read[:80]
        assert (int(g["trim5p"]), int(g["trim3p"])) == (w["trim5p"], w["trim3p"]), read[:80]
