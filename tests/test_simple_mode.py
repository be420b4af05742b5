"""CPU: simple mode (`--simple`, scanner_simple.py) and the sg_stats primitive.

* the C oracle reproduces the results the UNMODIFIED reference's BarcodeScannerSimple gave on the golden read set
  (tests/golden/golden_simple_v1.npz, made by tests/golden/make_golden_simple.py);
* qo_sg_stats agrees with qo_sg on score / end cell and its counters obey the alignment identities;
* the mirror's BarcodeScannerSimple (oracle-backed plan) equals the reference class on fresh reads, kit files included."""
import os
import sys

import numpy as np
import pytest

from tests import helpers

ROOT = helpers.ROOT
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)


def _cases():
    return [c["name"] for c in helpers.load_golden_simple()[1]]


@pytest.mark.parametrize("case_index", range(len(_cases())), ids=_cases())
def test_oracle_matches_reference_golden_simple(golden, case_index):
    data, _, _ = golden
    sdata, cases = helpers.load_golden_simple()
    case = cases[case_index]
    idx, want = sdata["idx_%d" % case_index], sdata["res_%d" % case_index]
    tables, sc = helpers.simple_tables_for_case(case)
    got = helpers.oracle_detect(tables, data["win5"][idx], data["tail3"][idx], data["wlen"][idx], data["read_len"][idx])
    helpers.assert_records_equal(got, want, case["name"])
    assert (want["layout"] == -1).all() and (want["barcode"] >= 0).sum() > 100


def test_sg_stats_counters():
    from qcat_b200 import config
    cfg = config.qcatConfig()
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for trial in range(300):
        n, m = int(rng.integers(1, 120)), int(rng.integers(1, 60))
        q = bytes(acgt[rng.integers(0, 5 if trial % 7 == 0 else 4, size=n)])
        r = bytes(acgt[rng.integers(0, 4, size=m)])
        if trial % 3 == 0:                                       # plant the reference inside the query: a real hit
            at = int(rng.integers(0, max(1, n - m + 1)))
            q = (q[:at] + r + q[at + m:])[:max(n, m)]
        for matrix, go, ge in ((cfg.matrix_barcode, 1, 1), (cfg.matrix, cfg.gap_open, cfg.gap_extend), (cfg.matrix_barcode, 3, 1)):
            sc, eq, er = helpers.oracle_sg(q, r, go, ge, matrix)
            sc2, eq2, er2, mt, sm, ln = helpers.oracle_sg_stats(q, r, go, ge, matrix)
            assert (sc, eq, er) == (sc2, eq2, er2)
            assert 0 <= mt <= sm + mt and mt <= ln and ln <= len(q) + len(r)
            assert mt <= min(len(q), len(r))
            if matrix is cfg.matrix_barcode and go == 1:
                # +1 / -1 / -1 scoring, no N in the template: score = matches - (length - matches) for ACGT queries
                if b"N" not in q:
                    assert sc == mt - (ln - mt)
    # exact copy: every base matches
    assert helpers.oracle_sg_stats(b"TTTACGTACGTAGG", b"ACGTACGTA", 1, 1, cfg.matrix_barcode)[3:] == (9, 9, 9)


@pytest.mark.skipif(not refloader.available(), reason="reference package not available")
@pytest.mark.parametrize("kit", ["standard", "extended", "file"])
def test_mirror_simple_scanner_equals_reference(tmp_path, kit):
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import scanner, synth
    from qcat_b200.tables import Tables
    if kit == "file":
        path = tmp_path / "barcodes.fasta"
        layouts = scanner.BarcodeScannerEPI2ME(kit="RBK004").layouts
        with open(path, "w") as fh:
            for b in layouts[0].barcode_set_1[:7]:
                fh.write(">%s some comment\n%s\n%s\n" % (b.name, b.sequence[:10], b.sequence[10:]))
        kit = str(path)
    ref = ref_scanner.factory(mode="simple", kit=kit)
    mine = scanner.factory(mode="simple", kit=kit)
    assert [(b.name, b.id, b.sequence) for b in mine.barcodes] == [(b.name, b.id, b.sequence) for b in ref.barcodes]
    assert mine.min_quality == ref.min_quality == 60 and mine.barcode_count() == ref.barcode_count()
    plan = helpers.OraclePlan(Tables.simple(mine.barcodes, ref_config.qcatConfig(), mine.min_quality))
    mine._plan_for = lambda qcat_config, layouts=None: plan
    src = scanner.BarcodeScannerEPI2ME(kit={"standard": "NBD103/NBD104", "extended": "PBC096"}.get(kit, "RBK004")).layouts
    reads = synth.windows_to_reads(synth.generate(src, 150, seed=19, sub=0.04, dele=0.03, ins=0.02)) + ["", "ACGT", "N" * 200]
    cfg = ref_config.qcatConfig()

    def key(r):
        b = r["barcode"]
        return (None if b is None else (b.name, b.id), r["barcode_score"], r["adapter"], r["adapter_end"], r["trim5p"],
                r["trim3p"], r["exit_status"])

    want = ref.detect_barcode_batch(reads, [None] * len(reads), cfg)
    got = mine.detect_barcode_batch(reads, [None] * len(reads), cfg)
    assert [key(r) for r in got] == [key(r) for r in want]
    assert sum(r["barcode"] is not None for r in want) > 20
    single = [mine.detect_barcode(r, None, cfg) for r in reads[:20]]
    assert [key(r) for r in single] == [key(ref.detect_barcode(r, None, cfg)) for r in reads[:20]]
    w = reads[3][:150]
    assert key(mine.scan(w, None, [], [], cfg)) == key(ref.scan(w, None, [], [], cfg))
    with pytest.raises(TypeError):
        scanner.factory(mode="simple")                          # kit=None: os.path.isfile(None), like the reference
