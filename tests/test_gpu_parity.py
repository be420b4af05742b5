"""GPU: the CUDA path (through the C ABI of libqcat_b200.so) is bit-identical to the golden vectors and to
the CPU oracle on seeded synthetic reads."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _case_ids():
    _, cases, _ = helpers.load_golden()
    return [c["name"] for c in cases]


@pytest.fixture(scope="module")
def engine():
    from qcat_b200 import engine as eng
    assert eng.device_count() > 0, "no CUDA device: the CUDA path has no CPU fallback"
    return eng


@pytest.mark.parametrize("force_generic", [False, True], ids=["auto", "generic"])
@pytest.mark.parametrize("case_index", range(len(_case_ids())), ids=_case_ids())
def test_cuda_matches_reference_golden(golden, engine, case_index, force_generic):
    data, cases, _ = golden
    case = cases[case_index]
    idx = data["idx_%d" % case_index]
    want = data["res_%d" % case_index]
    tables, sc = helpers.tables_for_case(case)
    plan = engine.DevicePlan(tables, device=0)
    plan.set_force_generic(force_generic)
    win5, tail3, wlen, read_len = (data[k][idx] for k in ("win5", "tail3", "wlen", "read_len"))
    subset = None
    if case["batch"]:
        names = [l.kit for l in sc.layouts]
        if len(set(names)) > 1:
            vote = plan.kit_vote(win5, tail3, wlen)
            np.testing.assert_array_equal(vote, helpers.oracle_kit_vote(tables, win5, tail3, wlen))
            subset = tables.kit_subset(helpers.kit_from_votes(vote, names))
    got = plan.detect(win5, tail3, wlen, read_len, subset)
    helpers.assert_records_equal(got, want, case["name"])
    plan.close()


import os

NBD196_FOLDER = os.path.join(helpers.ROOT, "qcat_b200", "resources", "nbd196")      # tools/make_nbd196.py (custom kit folder)
SYNTH = [("NBD103/NBD104", "epi2me", 20000), ("PBC096", "epi2me", 20000), ("NBD196", "epi2me", 12000), ("RBK004", "epi2me", 4000),
         ("RAB204/RAB214", "epi2me", 4000), (None, "epi2me", 6000), ("dual", "dual", 8000), ("DUAL", "epi2me", 4000)]


@pytest.mark.parametrize("force_generic", [False, True], ids=["auto", "generic"])
@pytest.mark.parametrize("kit,mode,n", SYNTH, ids=[str(s[0]) + "-" + s[1] for s in SYNTH])
def test_cuda_matches_oracle_on_synthetic(engine, kit, mode, n, force_generic):
    from qcat_b200 import config, scanner, synth
    from qcat_b200.tables import Tables
    cls = scanner.BarcodeScannerDual if mode == "dual" else scanner.BarcodeScannerEPI2ME
    sc = cls(kit=None if kit == "dual" else kit, kit_folder=NBD196_FOLDER if kit == "NBD196" else None)
    foreign = scanner.BarcodeScannerEPI2ME(kit="RBK001").layouts
    data = synth.generate(sc.layouts, n, seed=20261017 + n, foreign_layouts=foreign)
    tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)
    plan.set_force_generic(force_generic)
    got = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
    want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
    helpers.assert_records_equal(got, want, "%s/%s" % (kit, mode))
    assert (got["barcode"] >= 0).mean() > 0.3, "synthetic set should be mostly classifiable"
    plan.close()


def test_cuda_ragged_and_empty(engine):
    """Empty batch, empty reads, reads shorter than a barcode, windows of every length up to W."""
    from qcat_b200 import config, scanner
    from qcat_b200.tables import Tables, pack_windows
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)
    assert len(plan.detect_reads([])) == 0
    rng = np.random.default_rng(7)
    full = sc.layouts[0].get_adapter_sequences(sc.layouts[0].barcode_set_1[5].sequence)
    reads = ["", None, "A", "N" * 7]
    for n in range(1, 200):
        body = "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
        reads.append(body)
        reads.append((full + body)[:n + 40])
    packed = pack_windows(reads, 150)[:4]
    for force in (False, True):
        plan.set_force_generic(force)
        got = plan.detect(*packed)
        want = helpers.oracle_detect(tables, *packed)
        helpers.assert_records_equal(got, want, "ragged force_generic=%s" % force)
    plan.close()


def test_sg_batch_primitive_matches_oracle(engine):
    """qcb_sg_batch == oracle sg for both scoring schemes, affine gaps included, on ragged inputs."""
    from qcat_b200 import config
    cfg = config.qcatConfig()
    rng = np.random.default_rng(11)
    alphabet = "ACGTNacgtXR-"
    queries = ["".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n)) for n in (1, 2, 7, 24, 47, 150, 151, 400)]
    refs = ["".join("ACGTNX"[i] for i in rng.integers(0, 6, size=n)) for n in (1, 5, 39, 46, 102, 200)]
    for matrix, go, ge in ((cfg.matrix, 2, 2), (cfg.matrix_barcode, 1, 1), (cfg.matrix, 5, 1), (cfg.matrix_barcode, 3, 2)):
        score, end_query, end_ref = engine.sg_batch(queries, refs, go, ge, matrix, device=0)
        for qi, q in enumerate(queries):
            for ri, r in enumerate(refs):
                assert (score[qi, ri], end_query[qi, ri], end_ref[qi, ri]) == helpers.oracle_sg(q, r, go, ge, matrix), (q, r, go, ge)


def test_custom_config_uses_generic_kernels(engine):
    """A non-default qcatConfig (affine gaps, other scores, shorter window) still matches the oracle."""
    from qcat_b200 import config, scanner, synth
    from qcat_b200.tables import Tables
    cfg = config.qcatConfig()
    cfg.gap_open = 4
    cfg.gap_extend = 1
    cfg.match = 4
    cfg.mismatch = 3
    cfg.max_align_length = 120
    cfg.extracted_barcode_extension = 7
    cfg.barcode_context_length = 5
    sc = scanner.BarcodeScannerEPI2ME(kit="NBD104/NBD114")
    data = synth.generate(sc.layouts, 3000, seed=5, W=120, stride=128)
    tables = Tables(sc.layouts, cfg, "epi2me", sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)
    info = plan.info()
    assert info["fast_adapter"] == 0, "affine gaps (open != extend) must select the generic adapter kernel"
    assert info["fast_barcode"] == 1, "the barcode scheme is fixed at 1/1 (scanner_base.py:115-116): packed kernel"
    got = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
    want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
    helpers.assert_records_equal(got, want, "custom config")
    plan.close()
