"""CPU: host-side mirrors (config, layout, kit loading, window packing, vote rule) behave like the reference's."""
import os
import sys

import numpy as np
import pytest

from tests import helpers

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

needs_reference = pytest.mark.skipif(not refloader.available(), reason="reference package not available")


def test_get_placeholder_pos():
    """Same cases as the reference's test_get_placeholder (test_barcode.py:15-67)."""
    from qcat_b200.layout import AdapterLayout
    pos = AdapterLayout.get_placeholder_pos
    assert tuple(pos("NNNNN")) == (0, 4, 5)
    assert tuple(pos("AAAANNNNN")) == (4, 8, 5)
    assert tuple(pos("NNNNNAAAA")) == (0, 4, 5)
    assert tuple(pos("")) == (-1, -1, 0)
    assert tuple(pos("AATGTACTTCGTT")) == (-1, -1, 0)
    two = "AATGTACTTCGTTCAGTTACGTATTGCT" + "N" * 24 + "GTTTTCGCATTTATCGTG" + "N" * 10 + "AAACGC"
    assert tuple(pos(two, 0)) == (28, 51, 24)
    assert tuple(pos(two, 1)) == (70, 79, 10)


def test_scoring_matrices_match_the_reference_values():
    """Adapter matrix over ATGCNX with the N / X rows poked as in config.py:236-253; barcode matrix +-1."""
    from qcat_b200 import config
    cfg = config.qcatConfig()
    size, mat, mapper = config.matrix_arrays(cfg.matrix)
    m = mat.reshape(size, size)
    assert size == 7
    assert [m[i, i] for i in range(4)] == [5] * 4 and m[0, 1] == -2
    assert list(m[4, :5]) == [-1] * 5 and list(m[:4, 4]) == [-1] * 4          # N row / column
    assert list(m[5]) == [0] * 7 and list(m[:, 5]) == [0] * 7                  # X row / column
    assert list(m[6]) == [0] * 7                                               # wildcard
    assert mapper[ord("a")] == mapper[ord("A")] == 0 and mapper[ord("t")] == 1 and mapper[ord("?")] == 6
    size_b, mat_b, map_b = config.matrix_arrays(cfg.matrix_barcode)
    mb = mat_b.reshape(size_b, size_b)
    assert size_b == 6 and mb[4, 4] == 1 and mb[0, 4] == -1 and list(mb[5]) == [0] * 6
    cfg.match = -7            # setters normalise the sign like the reference (config.py:49-57)
    cfg.mismatch = 3
    assert cfg.match == 7 and cfg.mismatch == -3
    assert config.matrix_arrays(cfg.matrix)[1].reshape(7, 7)[1, 1] == 7


@needs_reference
def test_config_and_layout_mirrors_equal_the_reference_objects():
    refloader.load()
    from qcat import adapters as ref_adapters
    from qcat import config as ref_config
    from qcat_b200 import adapters, config
    a = config.matrix_arrays(config.qcatConfig().matrix)
    b = config.matrix_arrays(ref_config.qcatConfig().matrix)
    assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    a = config.matrix_arrays(config.qcatConfig().matrix_barcode)
    b = config.matrix_arrays(ref_config.qcatConfig().matrix_barcode)
    assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    mine = {(l.kit, l.sequence): l for l in adapters.populate_adapter_layouts()}
    for ref in ref_adapters.populate_adapter_layouts():
        l = mine[(ref.kit, ref.sequence)]
        for k in (0, 1):
            assert l.get_barcode_end(k) == ref.get_barcode_end(k)
            assert l.get_barcode_length(k) == ref.get_barcode_length(k)
            assert l.get_upstream_context(11, k) == ref.get_upstream_context(11, k)
            assert l.get_downstream_context(11, k) == ref.get_downstream_context(11, k)
        assert l.is_double_barcode() == ref.is_double_barcode() and l.trim_offset == ref.trim_offset
        assert l.get_adapter_length() == ref.get_adapter_length() and l.auto_detect == ref.auto_detect


def test_kit_selection_and_factory():
    from qcat_b200 import scanner
    assert len(scanner.BarcodeScannerEPI2ME().layouts) == 12                     # auto_detect layouts
    assert len(scanner.BarcodeScannerEPI2ME(kit="Auto").layouts) == 12
    assert [l.kit for l in scanner.BarcodeScannerEPI2ME(kit="pbc096").layouts] == ["PBC096", "PBC096"]
    assert scanner.BarcodeScannerEPI2ME(kit="nonexistent").layouts == []
    assert scanner.factory(mode="dual").min_quality == 60 and scanner.factory().min_quality == 58
    assert scanner.factory(mode="guppy").get_name() == "epi2me"
    assert sorted(scanner.get_modes()) == ["dual", "epi2me", "simple"]
    assert "PBC096" in scanner.get_kits() and scanner.get_kits()[0] == "Auto"
    simple = scanner.factory(mode="simple", kit="standard")
    assert simple.get_name() == "simple" and simple.min_quality == 60 and len(simple.barcodes) == 24
    assert len(scanner.factory(mode="simple", kit="extended").barcodes) == 120
    with pytest.raises(RuntimeError):
        scanner.factory(mode="brill")


def test_pack_windows_is_extract_align_sequence():
    """win5 = read[:W], tail3 = read[-W:] (scanner_base.py:223-244), ragged and empty reads included."""
    from qcat_b200.tables import pack_windows
    from qcat_b200.scanner import revcomp
    reads = ["", None, "ACG", "A" * 150, "ACGT" * 100, "acgtRYN-" * 30]
    win5, tail3, wlen, read_len, stride = pack_windows(reads, 150)
    assert stride == 160 and win5.shape == (6, 160)
    for i, r in enumerate(reads):
        r = r or ""
        assert read_len[i] == len(r) and wlen[i] == min(len(r), 150)
        assert bytes(win5[i, :wlen[i]]).decode() == r[:150]
        assert bytes(tail3[i, :wlen[i]]).decode() == r[-150:]
    assert revcomp("ACGTNacgtRYKMx") == "xKMRYacgtNACGT"


def test_kit_vote_tie_rule():
    """Most abundant kit, first seen wins ties (dict order + stable sort, scanner_base.py:645-660)."""
    names = ["A", "B", "B", "C"]
    vote = np.array([3, 1, 0, 2, 3, 0], dtype=np.int32)       # C, B, A, B, C, A -> all 2: first seen is C
    assert helpers.kit_from_votes(vote, names) == "C"
    assert helpers.kit_from_votes(np.array([1, 2, 0], dtype=np.int32), names) == "B"
    assert helpers.kit_from_votes(np.array([], dtype=np.int32), names) is None


def test_filter_barcodes_and_counts():
    from qcat_b200 import scanner
    from qcat_b200.adapters import Barcode
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096", enable_filter_barcodes=True)
    b1, b2 = Barcode("barcode01", 1, "A", True), Barcode("barcode02", 2, "C", True)
    results = [scanner.build_return_dict(b1, 90.0, sc.layouts[0], 50, 0) for _ in range(100)]
    results += [scanner.build_return_dict(b2, 90.0, sc.layouts[0], 50, 0) for _ in range(5)]
    results += [scanner.empty_return_dict()]
    counts = {}
    for r in results:
        sc.update_barcode_count(r, counts)
    assert counts == {1: 100, 2: 5, "0": 1}
    out = sc.filter_barcodes(counts, list(results))
    assert sum(1 for r in out if r["barcode"] is b1) == 100
    assert all(r["barcode"] is None for r in out[100:])       # 5 <= int(100 * 0.05): dropped


def test_synthetic_generator_is_deterministic_and_well_formed():
    from qcat_b200 import scanner, synth
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    a = synth.generate(sc.layouts, 500, seed=3)
    b = synth.generate(sc.layouts, 500, seed=3)
    for k in ("win5", "tail3", "wlen", "read_len", "truth_barcode"):
        assert (a[k] == b[k]).all()
    assert a["win5"].shape == (500, 160) and (a["wlen"] == 150).all() and (a["read_len"] >= 300).all()
    assert set(np.unique(a["win5"][:, :150])) <= set(b"ACGTN")
    assert (a["win5"][:, 150:] == 0).all()
    reads = synth.windows_to_reads(a, range(5))
    assert [len(r) for r in reads] == list(a["read_len"][:5])


def test_oracle_counts_reference_cells():
    """Algorithmic DP cells per read for PBC096 on un-barcoded windows = 35 700 + 1 209 600."""
    from qcat_b200 import config, scanner
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    rng = np.random.default_rng(1)
    win = np.zeros((50, 160), dtype=np.uint8)
    win[:, :150] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(50, 150))]
    wlen = np.full(50, 150, dtype=np.int32)
    cells, full = helpers.oracle_count_cells(tables, win, win, wlen)
    assert full == 100 and cells == 50 * (35700 + 1209600)


@pytest.mark.parametrize("mode", ["epi2me", "dual"])
def test_batch_record_conversion_equals_per_record(mode):
    """GpuScannerMixin._records_to_dicts (columns converted once) builds the same dicts, with the same Barcode /
    AdapterLayout objects, as _record_to_dict row by row."""
    from qcat_b200 import _ffi, config, scanner
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096") if mode == "epi2me" else scanner.BarcodeScannerDual()
    tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)

    class FakePlan(object):
        pass
    plan = FakePlan()
    plan.tables = tables
    rng = np.random.default_rng(3)
    recs = np.zeros(500, dtype=_ffi.RESULT_DTYPE)
    recs["layout"] = rng.integers(-1, len(sc.layouts), 500)
    recs["barcode"] = rng.integers(0, 96, 500)
    recs["barcode_score"] = rng.random(500) * 100
    recs["adapter_end"] = rng.integers(0, 150, 500)
    recs["trim5p"] = rng.integers(0, 150, 500)
    recs["trim3p"] = rng.integers(150, 9000, 500)
    recs["exit_status"] = np.where(recs["layout"] < 0, rng.choice([1, 1002], 500), 0)
    one_by_one = [sc._record_to_dict(plan, r) for r in recs]
    batch = sc._records_to_dicts(plan, recs)
    assert batch == one_by_one
    for a, b in zip(batch, one_by_one):
        assert a["adapter"] is b["adapter"] and type(a["barcode_score"]) is float and type(a["trim5p"]) is int
        if mode == "epi2me":
            assert a["barcode"] is b["barcode"]


@needs_reference
def test_custom_kit_folder_nbd196_matches_the_reference():
    """BASELINE configs[2] names EXP-NBD196, which qcat 1.1.0 does not ship: the synthetic kit of tools/make_nbd196.py
    is loaded through kit_folder by the reference (adapters.py:138-162) and by the mirror alike, and the oracle agrees
    with the reference's Python on reads of that kit (batch and single-read API)."""
    import os
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import config, scanner, synth
    from qcat_b200.tables import Tables, pack_windows
    folder = os.path.join(helpers.ROOT, "qcat_b200", "resources", "nbd196")
    ref = ref_scanner.factory(kit="NBD196", kit_folder=folder)
    mine = scanner.factory(kit="NBD196", kit_folder=folder)
    assert sorted((l.kit, l.sequence, len(l.barcode_set_1)) for l in mine.layouts) == \
        sorted((l.kit, l.sequence, len(l.barcode_set_1)) for l in ref.layouts)
    assert len(mine.layouts) == 2 and all(len(l.barcode_set_1) == 96 for l in mine.layouts)
    # same layout order as the reference object for the comparison (the folder glob is unsorted on both sides)
    order = {l.sequence: i for i, l in enumerate(ref.layouts)}
    mine.layouts.sort(key=lambda l: order[l.sequence])
    data = synth.generate(mine.layouts, 160, seed=9, mean_len=900.0)
    reads = synth.windows_to_reads(data) + ["", "ACGT", "N" * 400]
    cfg = ref_config.qcatConfig()
    want = ref.detect_barcode_batch(reads, [None] * len(reads), cfg)
    tables = Tables(mine.layouts, config.qcatConfig(), "epi2me", mine.min_quality)
    win5, tail3, wlen, read_len, _ = pack_windows(reads, 150)
    got = helpers.oracle_detect(tables, win5, tail3, wlen, read_len)
    assert sum(r["barcode"] is not None for r in want) > 100
    assert len({r["barcode"].id for r in want if r["barcode"]}) > 40          # ids well beyond the 12 of NBD104
    for g, w in zip(got, want):
        if w["barcode"] is None:
            assert g["barcode"] < 0 and g["exit_status"] == w["exit_status"]
        else:
            layout = mine.layouts[int(g["layout"])]
            assert layout.sequence == w["adapter"].sequence
            assert layout.barcode_set_1[int(g["barcode"])].id == w["barcode"].id
            assert float(g["barcode_score"]) == w["barcode_score"] and int(g["adapter_end"]) == w["adapter_end"]
        assert (int(g["trim5p"]), int(g["trim3p"])) == (w["trim5p"], w["trim3p"])


def test_vectorised_batch_kit_vote_equals_reference_rule():
    """fastx._batch_kits (numpy, per CLI batch) picks the same kit as the reference's dict-insertion-order + stable-sort
    rule (scanner_base.py:657-678, mirrored by GpuScannerMixin._kit_from_votes), ties included."""
    from qcat_b200 import fastx
    from qcat_b200.scanner import GpuScannerMixin
    rng = np.random.default_rng(11)
    names = ["A", "A", "B", "C", "C", "C", "D"]                         # layouts -> kits
    kit_names = list(dict.fromkeys(names))
    kit_of_layout = np.array([kit_names.index(k) for k in names], dtype=np.int64)
    for trial in range(200):
        n = int(rng.integers(1, 60))
        batch = int(rng.integers(1, 12))
        vote = rng.integers(0, len(names), size=n).astype(np.int32)
        if trial % 3 == 0:                                               # force ties
            vote = np.resize(rng.permutation(len(names)), n).astype(np.int32)
        got = fastx._batch_kits(vote, kit_of_layout, len(kit_names), batch)
        want = [GpuScannerMixin._kit_from_votes(vote[lo:lo + batch], names) for lo in range(0, n, batch)]
        assert [kit_names[k] for k in got] == want


def test_vectorised_filter_barcodes_equals_reference_rule():
    """fastx._filter_barcodes == BarcodeScanner.filter_barcodes per batch (scanner_base.py:680-712): ids seen in
    <= int(5 % of the most frequent key) reads are emptied, 'none' counts as a key, trims are reset."""
    from qcat_b200 import _ffi, config, fastx, scanner
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)

    class FakePlan(object):
        pass
    plan = FakePlan()
    plan.tables = tables
    rng = np.random.default_rng(5)
    n, batch = 900, 300
    recs = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    weights = np.r_[np.full(6, 0.15), np.full(90, 0.1 / 90)]            # six frequent barcodes, a long tail of rare ones
    recs["barcode"] = rng.choice(96, size=n, p=weights / weights.sum())
    recs["layout"] = rng.integers(0, 2, size=n)
    none = rng.random(n) < 0.2
    recs["layout"][none] = -1
    recs["barcode"][none] = -1
    recs["exit_status"] = np.where(none, 1, 0)
    recs["barcode_score"] = np.where(none, 0.0, 77.0)
    recs["trim5p"] = 30
    recs["trim3p"] = 500
    got = recs.copy()
    fastx._filter_barcodes(tables, got, batch)
    want = []
    for lo in range(0, n, batch):
        dicts = sc._records_to_dicts(plan, recs[lo:lo + batch])
        count = {}
        for d in dicts:
            sc.update_barcode_count(d, count)
        want += sc.filter_barcodes(count, dicts)
    assert sum(w["barcode"] is None for w in want) > none.sum()          # something was filtered
    for g, w in zip(got, want):
        assert (g["barcode"] < 0) == (w["barcode"] is None)
        assert (int(g["trim5p"]), int(g["trim3p"]), int(g["exit_status"])) == (w["trim5p"], w["trim3p"], w["exit_status"])


def test_batch_straddler_equals_plain_batching():
    """fastx._BatchStraddler: chunks that end anywhere still give every read the result its CLI batch (consecutive
    groups of `batch_size` reads, cli.py:500-513) would give it, chunks leave complete and in input order."""
    from qcat_b200 import _ffi, fastx

    class FakeChunk(object):
        def __init__(self, ident):
            self.ident, self.released = ident, False

        def release(self):
            self.released = True

    rng = np.random.default_rng(12)
    for trial in range(30):
        batch_size = int(rng.integers(1, 40))
        sizes = [int(v) for v in rng.integers(0, 3 * batch_size, size=int(rng.integers(1, 25)))]
        total = sum(sizes)
        read_ids = np.arange(total, dtype=np.int64)
        calls = []

        def score(packed):
            # a "score" that depends on the composition of the batch a read is in: (first read of its batch, batch size)
            ids = packed[3]
            calls.append(len(ids))
            out = np.zeros(len(ids), dtype=_ffi.RESULT_DTYPE)
            for lo in range(0, len(ids), batch_size):
                part = ids[lo:lo + batch_size]
                out["trim5p"][lo:lo + len(part)] = part[0]
                out["trim3p"][lo:lo + len(part)] = len(part)
                out["adapter_end"][lo:lo + len(part)] = part - part[0]
            return out

        straddler = fastx._BatchStraddler(batch_size)
        done, pos = [], 0
        for ident, n in enumerate(sizes):
            ids = read_ids[pos:pos + n]
            pos += n
            packed = (np.zeros((n, 16), np.uint8), np.zeros((n, 16), np.uint8), np.zeros(n, np.int32), ids)
            done += straddler.add(FakeChunk(ident), packed, score)
        done += straddler.finish(score)
        straddler.release_all(failed=False)
        assert [c.ident for c, _, _ in done] == list(range(len(sizes)))
        assert all(n % batch_size == 0 for n in calls[:-1]) or not calls
        got = np.concatenate([r for _, _, r in done]) if done else np.zeros(0, dtype=_ffi.RESULT_DTYPE)
        assert len(got) == total
        np.testing.assert_array_equal(got["trim5p"], read_ids // batch_size * batch_size)
        np.testing.assert_array_equal(got["adapter_end"], read_ids % batch_size)
        np.testing.assert_array_equal(got["trim3p"], np.minimum(batch_size, total - read_ids // batch_size * batch_size))
        for (chunk, read_len, results), n in zip(done, sizes):
            assert len(results) == n and len(read_len) == n and not chunk.released
