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


@pytest.mark.parametrize("mode,kit", [("epi2me", "PBC096"), ("epi2me", "VMK001"), ("epi2me", None), ("dual", None)])
def test_long_window_row_chunks_equal_single_thread_scan(mode, kit):
    """qcb_scan on long windows (--detect-middle bodies): the row-chunked adapter stage (k_adapter_long, 1024-row chunks
    with a provably sufficient warm-up) gives exactly the records of the one-thread-per-template generic kernel, with
    adapters planted at and across chunk boundaries, at the window ends, and in windows of every chunk-count."""
    from qcat_b200 import config, engine, scanner
    from qcat_b200.tables import Tables
    rng = np.random.default_rng(17)
    sc = scanner.factory(mode=mode, kit=kit, device=0)
    tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)

    def noisy(seq):
        out = []
        for ch in seq:
            r = rng.random()
            if r < 0.03:
                continue
            out.append("ACGT"[rng.integers(4)] if r < 0.07 else ch)
            if rng.random() < 0.02:
                out.append("ACGT"[rng.integers(4)])
        return "".join(out)

    windows = []
    for length in [0, 1, 500, 1023, 1024, 1025, 2047, 2048, 2049, 3000, 4096, 5000, 7777]:
        for rep in range(6):
            body = list("".join("ACGT"[i] for i in rng.integers(0, 4, size=length)))
            if length > 200 and rep:
                layout = sc.layouts[int(rng.integers(len(sc.layouts)))]
                bc = layout.barcode_set_1[int(rng.integers(len(layout.barcode_set_1)))].sequence
                adapter = layout.get_adapter_sequences(bc)
                if layout.is_double_barcode():
                    bc2 = layout.barcode_set_2[int(rng.integers(len(layout.barcode_set_2)))].sequence
                    p2 = layout.barcode_pos_2
                    adapter = adapter[:p2.start] + bc2 + adapter[p2.end + 1:]
                adapter = noisy(adapter) if rep > 2 else adapter
                # around every chunk boundary, at both ends, and somewhere random
                anchors = [1024 * k for k in range(1, length // 1024 + 1)] + [0, length - len(adapter), int(rng.integers(0, length))]
                pos = int(anchors[int(rng.integers(len(anchors)))] - rng.integers(0, len(adapter) + 1))
                pos = max(0, min(pos, length - len(adapter)))
                body[pos:pos + len(adapter)] = list(adapter)
                if rep == 5:                                        # a second copy: ties between far-apart rows
                    pos2 = max(0, min(int(rng.integers(0, length)), length - len(adapter)))
                    body[pos2:pos2 + len(adapter)] = list(adapter)
            windows.append("".join(body)[:length] if length else "")
    subset = list(range(len(sc.layouts)))
    plan.set_force_generic(1)
    chunked = plan.scan_windows(windows, subset)
    plan.set_force_generic(2)
    serial = plan.scan_windows(windows, subset)
    plan.set_force_generic(0)
    helpers.assert_records_equal(chunked, serial, "row-chunked vs single-thread scan")
    assert (serial["layout"] >= 0).sum() > len(windows) // 3
    plan.close()


def test_plans_of_different_kits_coexist():
    """Several plans in one process (the scanner caches up to four): the packed kernels' dynamic shared-memory opt-in is
    a per-device maximum shared by all of them, so creating a small-kit plan must not break a large-kit one."""
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    made = []
    for mode, kit in (("epi2me", "PBC096"), ("epi2me", "NBD103/NBD104"), ("dual", None), ("epi2me", None)):
        sc = scanner.factory(mode=mode, kit=kit, device=0)
        tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
        data = synth.generate(sc.layouts if (kit or mode == "dual") else scanner.factory(kit="RBK004").layouts, 700, seed=5)
        made.append((engine.DevicePlan(tables, device=0), tables, data))
    for _ in range(2):
        for plan, tables, data in made + made[::-1]:
            got = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
            want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
            helpers.assert_records_equal(got, want, "interleaved plans")
    for plan, _, _ in made:
        plan.close()


def test_host_entry_points_reject_bad_window_lengths():
    """Window lengths are validated chunk by chunk on the host path: a bad entry deep inside a large call fails the call
    with its index, and the plan stays usable."""
    from qcat_b200 import _ffi, config, engine, scanner, synth
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="NBD103/NBD104")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    data = synth.generate(sc.layouts, 150000, seed=8)
    plan = engine.DevicePlan(tables, device=0)
    try:
        good = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        for index, value in ((140000, 151), (7, -1)):
            wlen = data["wlen"].copy()
            wlen[index] = value
            with pytest.raises(_ffi.QcbError, match=r"wlen\[%d\]" % index):
                plan.detect(data["win5"], data["tail3"], wlen, data["read_len"])
        again = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        helpers.assert_records_equal(again, good, "plan reusable after a rejected call")
    finally:
        plan.close()
