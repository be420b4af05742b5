"""GPU, BASELINE-size batches (1 M reads): size-independent properties of the CUDA path.

The oracle is too slow to check a million reads, so at full size the checks are: the packed kernels agree with
the independent generic kernels on every record; a random 20 k subset agrees with the CPU oracle; results do not
depend on batch order or chunk boundaries; the device histogram accounts for every read."""
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


def test_random_subset_matches_oracle(full_batch):
    data, tables, plan, fast = full_batch
    idx = np.sort(np.random.default_rng(5).choice(N_FULL, size=20000, replace=False))
    want = helpers.oracle_detect(tables, data["win5"][idx], data["tail3"][idx], data["wlen"][idx], data["read_len"][idx])
    helpers.assert_records_equal(fast[idx], want, "1M batch, 20k subset vs oracle")


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


def test_dual_and_small_kits_at_scale():
    """configs[1] (12 barcodes) and configs[3] (dual 24 x 96): packed == generic on 200 k reads each."""
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    for cls, kit, mode in ((scanner.BarcodeScannerEPI2ME, "NBD103/NBD104", "epi2me"), (scanner.BarcodeScannerDual, None, "dual")):
        sc = cls(kit=kit)
        data = synth.generate(sc.layouts, 200000, seed=77)
        tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
        plan = engine.DevicePlan(tables, device=0)
        fast = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        plan.set_force_generic(True)
        generic = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        helpers.assert_records_equal(fast, generic, "packed vs generic, %s" % mode)
        assert (fast["barcode"] >= 0).mean() > 0.3
        plan.close()
