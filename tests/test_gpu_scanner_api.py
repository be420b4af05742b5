"""GPU: the reference-shaped Python API (detect_barcode / detect_barcode_batch / scan) over the CUDA plan.
Reads like the reference's own tests (qcat/test/test_barcode.py) where those pin results."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _reads_from_golden(data, idx):
    """Rebuild read strings from the stored windows (the detection path sees only windows + length)."""
    reads = []
    for i in idx:
        n = int(data["read_len"][i])
        k = int(data["wlen"][i])
        head = bytes(data["win5"][i, :k]).decode("latin-1")
        tail = bytes(data["tail3"][i, :k]).decode("latin-1")
        if n <= 150:
            reads.append(head)
        elif n < 300:
            reads.append(head + tail[-(n - 150):])
        else:
            reads.append(head + "A" * (n - 300) + tail)
        assert len(reads[-1]) == n
    return reads


def _encode(sc, result):
    layout = -1 if result["adapter"] is None else [i for i, l in enumerate(sc.layouts) if l is result["adapter"]][0]
    barcode = -1
    if result["barcode"] is not None:
        L = result["adapter"]
        if sc.get_name() == "dual":
            a, b = result["barcode"].id.split("/")
            i1 = [i for i, bc in enumerate(L.barcode_set_1) if str(bc.id) == a][0]
            i2 = [i for i, bc in enumerate(L.barcode_set_2) if str(bc.id) == b][0]
            barcode = i1 * len(L.barcode_set_2) + i2
        else:
            barcode = [i for i, bc in enumerate(L.barcode_set_1) if bc is result["barcode"]][0]
    return (layout, barcode, float(result["barcode_score"]), result["adapter_end"], result["trim5p"], result["trim3p"],
            result["exit_status"])


@pytest.mark.parametrize("name", ["auto/batch/barcode_1k.fastq", "auto/batch/pbk004.fastq", "dual/batch/small",
                                  "PBC096/batch/nobarcode_1k", "auto/batch/adversarial"])
def test_detect_barcode_batch_matches_golden(golden, name):
    from qcat_b200 import _ffi
    data, cases, _ = golden
    ci = [i for i, c in enumerate(cases) if c["name"] == name][0]
    idx = data["idx_%d" % ci]
    sc = helpers.scanner_for_case(cases[ci], device=0)
    reads = _reads_from_golden(data, idx)
    results = sc.detect_barcode_batch(reads, [None] * len(reads))
    got = np.array([_encode(sc, r) for r in results], dtype=_ffi.RESULT_DTYPE)
    helpers.assert_records_equal(got, data["res_%d" % ci], name)
    assert sc.override_kit_name is None


def test_detect_barcode_single_matches_golden(golden):
    from qcat_b200 import _ffi
    data, cases, ranges = golden
    ci = 0                                            # auto/single/all
    sc = helpers.scanner_for_case(cases[ci], device=0)
    idx = list(range(*ranges["pbk004.fastq"])) + list(range(*ranges["literals"])) + list(range(*ranges["adversarial"]))[:60]
    reads = _reads_from_golden(data, idx)
    got = np.array([_encode(sc, sc.detect_barcode(r)) for r in reads], dtype=_ffi.RESULT_DTYPE)
    helpers.assert_records_equal(got, data["res_%d" % ci][idx], "auto/single")


def test_reference_known_answers():
    """Barcode-name KATs of the reference test-suite (test_barcode.py:70-85, :307-322)."""
    from qcat_b200.scanner import BarcodeScannerEPI2ME, factory
    data, cases, ranges = helpers.load_golden()
    lit = ranges["literals"][0]
    reads = _reads_from_golden(data, [lit, lit + 1, lit + 2, lit + 3])
    sc = BarcodeScannerEPI2ME(device=0)
    assert sc.detect_barcode(reads[0])["barcode"].name == "barcode02"
    assert sc.detect_barcode(reads[1])["barcode"].name == "barcode03"
    assert sc.detect_barcode(reads[2])["barcode"].name == "barcode03"
    assert sc.detect_barcode("")["barcode"] is None
    assert sc.detect_barcode(None)["barcode"] is None
    dual = factory(mode="dual", device=0)
    r = dual.detect_barcode(reads[3])
    assert r["barcode"].name == "barcode06/95" and r["barcode"].id == "6/95"
    assert (r["adapter_end"], r["trim5p"], r["trim3p"]) == (120, 120, 1688)
    assert r["barcode_score"] == 84.78260869565217
    with pytest.raises(RuntimeError):
        factory(mode="nonsense")


def test_scan_single_window_and_truncating_zip():
    from qcat_b200.scanner import BarcodeScannerEPI2ME
    from qcat_b200 import config
    sc = BarcodeScannerEPI2ME(kit="RBK001", device=0)
    layout = sc.layouts[0]
    window = layout.get_adapter_sequences(layout.barcode_set_1[2].sequence) + "ACGTTGCA" * 10
    d = sc.scan(window, None, sc.layouts, [], config.qcatConfig())
    assert d["barcode"] is layout.barcode_set_1[2] and d["adapter"] is layout
    assert d["barcode_score"] == 100.0 and d["adapter_end"] == len(layout.sequence) - 1
    assert d["trim5p"] == 0 and d["trim3p"] == 0 and d["exit_status"] == 0
    # detect_barcode_batch(read_qualities=[None]) silently truncates to one read (scanner_base.py:714, :723)
    assert len(sc.detect_barcode_batch([window, window, window])) == 1
    assert sc.detect_barcode_batch([]) == []
