"""GPU, BASELINE-size batches (1 M reads): bit-identity with the CPU oracle on every read, plus size-independent
properties of the CUDA path.

The 1 M-read batch is drawn (with repetition) from 250 000 distinct synthetic reads plus 5 000 ragged ones, so the
oracle only has to score the distinct reads (~25 s on the box's cores) for all one million records to be compared.
Further checks: the packed kernels agree with the independent generic kernels on every record; results do not depend
on batch order or chunk boundaries; the device histogram accounts for every read."""
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

N_FULL = 1000000


@pytest.fixture(scope="module")
def full_batch():
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    foreign = scanner.BarcodeScannerEPI2ME(kit="RBK001").layouts
    unique = synth.generate(sc.layouts, 250000, seed=20261019, foreign_layouts=foreign)
    rng = np.random.default_rng(3)
    pick = rng.integers(0, 250000, size=N_FULL)
    data = {k: np.ascontiguousarray(unique[k][pick]) for k in ("win5", "tail3", "wlen", "read_len")}
    # ragged tail: a few thousand short / empty reads mixed in
    short = rng.integers(0, N_FULL, size=5000)
    newlen = rng.integers(0, 150, size=5000).astype(np.int32)
    data["wlen"][short] = newlen
    data["read_len"][short] = newlen
    for i, n in zip(short, newlen):
        data["tail3"][i, :n] = data["win5"][i, :n]
        data["win5"][i, n:] = 0
        data["tail3"][i, n:] = 0
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)
    result = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
    data["unique"], data["pick"], data["short"] = unique, pick, np.unique(short)
    yield data, tables, plan, result
    plan.close()


def test_packed_and_generic_kernels_agree_on_a_million_reads(full_batch):
    data, tables, plan, fast = full_batch
    assert plan.info()["fast_adapter"] == 1 and plan.info()["fast_barcode"] == 1
    plan.set_force_generic(True)
    generic = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
    plan.set_force_generic(False)
    helpers.assert_records_equal(fast, generic, "packed vs generic kernels, 1M reads")
    called = (fast["barcode"] >= 0).mean()
    assert 0.5 < called < 0.95


def test_every_read_of_the_million_matches_oracle(full_batch):
    """bit-identical records on 1 M synthetic reads of the headline config.  The oracle scores the
    250 000 distinct reads and the ragged ones; every one of the 1 M device records is compared with its read's."""
    data, tables, plan, fast = full_batch
    u = data["unique"]
    want = helpers.oracle_detect(tables, u["win5"], u["tail3"], u["wlen"], u["read_len"])[data["pick"]]
    idx = data["short"]
    want[idx] = helpers.oracle_detect(tables, data["win5"][idx], data["tail3"][idx], data["wlen"][idx], data["read_len"][idx])
    helpers.assert_records_equal(fast, want, "1M batch, every record vs oracle")


def test_order_and_chunking_invariance(full_batch):
    data, tables, plan, fast = full_batch
    perm = np.random.default_rng(9).permutation(N_FULL)[:300000]
    shuffled = plan.detect(data["win5"][perm], data["tail3"][perm], data["wlen"][perm], data["read_len"][perm])
    helpers.assert_records_equal(shuffled, fast[perm], "permutation invariance")
    cut = 123457                                   # not a multiple of any tile / chunk size
    a = plan.detect(data["win5"][:cut], data["tail3"][:cut], data["wlen"][:cut], data["read_len"][:cut])
    b = plan.detect(data["win5"][cut:400000], data["tail3"][cut:400000], data["wlen"][cut:400000], data["read_len"][cut:400000])
    helpers.assert_records_equal(np.concatenate([a, b]), fast[:400000], "chunk boundary invariance")


def test_device_histogram_accounts_for_every_read(full_batch):
    import torch
    from qcat_b200 import dist as qdist
    data, tables, plan, fast = full_batch
    base, n_bins = plan.histogram_layout()
    d_res = torch.from_numpy(fast.view(np.uint8).reshape(-1).copy()).cuda()
    d_counts = torch.zeros(n_bins, dtype=torch.int64, device="cuda")
    plan.histogram_device(d_res.data_ptr(), N_FULL, base, d_counts.data_ptr(), n_bins,
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    counts = d_counts.cpu().numpy()
    assert counts.sum() == N_FULL
    np.testing.assert_array_equal(counts, qdist.histogram_bins(fast, base, n_bins))
    hist = qdist.barcode_histogram(tables, counts, base)
    assert hist["none"] == int((fast["barcode"] < 0).sum()) and len(hist) >= 90


NBD196_FOLDER = os.path.join(helpers.ROOT, "qcat_b200", "resources", "nbd196")


@pytest.mark.parametrize("mode,kit,kit_folder,n", [("epi2me", "NBD103/NBD104", None, 200000), ("dual", None, None, 200000),
                                                   ("epi2me", "NBD196", NBD196_FOLDER, 100000)])
def test_other_configs_at_scale_match_oracle(mode, kit, kit_folder, n):
    """configs[1] (12 barcodes), configs[3] (dual 24 x 96) and the synthetic EXP-NBD196 kit: every record of 100-200 k
    reads equals the CPU oracle's, and the packed kernels equal the generic ones."""
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    cls = scanner.BarcodeScannerDual if mode == "dual" else scanner.BarcodeScannerEPI2ME
    sc = cls(kit=kit, kit_folder=kit_folder)
    data = synth.generate(sc.layouts, n, seed=77)
    tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
    plan = engine.DevicePlan(tables, device=0)
    try:
        fast = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        assert plan.info()["fast_adapter"] == 1 and plan.info()["fast_barcode"] == 1
        plan.set_force_generic(True)
        generic = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        helpers.assert_records_equal(fast, generic, "packed vs generic, %s %s" % (mode, kit))
        want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
        helpers.assert_records_equal(fast, want, "%d reads vs oracle, %s %s" % (n, mode, kit))
        assert (fast["barcode"] >= 0).mean() > 0.3
    finally:
        plan.close()
