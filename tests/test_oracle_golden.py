"""CPU: the C oracle (oracle/qcat_oracle.c) reproduces the reference's Python results stored in tests/golden."""
import pytest

from tests import helpers


def _case_ids():
    _, cases, _ = helpers.load_golden()
    return [c["name"] for c in cases]


@pytest.mark.parametrize("case_index", range(len(_case_ids())), ids=_case_ids())
def test_oracle_matches_reference_golden(golden, case_index):
    data, cases, _ = golden
    case = cases[case_index]
    idx = data["idx_%d" % case_index]
    want = data["res_%d" % case_index]
    tables, sc = helpers.tables_for_case(case)
    win5, tail3, wlen, read_len = (data[k][idx] for k in ("win5", "tail3", "wlen", "read_len"))
    subset = None
    if case["batch"]:
        names = [l.kit for l in sc.layouts]
        if len(set(names)) > 1:
            vote = helpers.oracle_kit_vote(tables, win5, tail3, wlen)
            kit = helpers.kit_from_votes(vote, names)
            subset = tables.kit_subset(kit)
    got = helpers.oracle_detect(tables, win5, tail3, wlen, read_len, subset)
    helpers.assert_records_equal(got, want, case["name"])


def test_known_answer_end_query_101():
    """The reference's one pinned end position (test_barcode.py:291-304): RBK001 on read_bc3_exact -> 101."""
    from qcat_b200 import config, scanner
    data, cases, ranges = helpers.load_golden()
    lay = scanner.get_adapter_by_name("RBK001")
    assert len(lay) == 1
    cfg = config.qcatConfig()
    # the golden file only keeps 150-nt windows; the alignment ends at 101, well inside the window
    i = ranges["literals"][0] + 1
    window = bytes(data["win5"][i, :150])
    sc, eq, er = helpers.oracle_sg(window, lay[0].get_adapter_sequences(), cfg.gap_open, cfg.gap_extend, cfg.matrix)
    assert eq == 101
    assert sc == 57 * 5 - 24       # 57 matching adapter bases, 24 N-masked barcode bases at -1 each
