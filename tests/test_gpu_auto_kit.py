"""GPU: the one-pass auto-kit flow (qcb_detect_auto, `-k auto` = the CLI default; scanner_base.py:618-678, :714-733).

One device pass -- adapter stage over all layouts, per-batch kit election on the device, kit-restricted selection on
the adapter scores already computed -- must equal the reference's two steps: detect_kit over the batch, then
detect_barcode per read restricted to the elected kit (CPU oracle: qo_kit_vote + get_most_abundant_kits + qo_detect)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _mixed_batches(n_batches, batch_size, seed, kits=("PBC096", "RBK004", "NBD103/NBD104", "RAB204/RAB214")):
    """Reads of `n_batches` consecutive batches; every batch is dominated by one kit (70 %) with reads of the others and
    short / empty reads mixed in."""
    from qcat_b200 import scanner, synth
    rng = np.random.default_rng(seed)
    pools = {k: synth.generate(scanner.BarcodeScannerEPI2ME(kit=k).layouts, n_batches * batch_size, seed=seed + i)
             for i, k in enumerate(kits)}
    parts = {f: [] for f in ("win5", "tail3", "wlen", "read_len")}
    for b in range(n_batches):
        main = kits[b % len(kits)]
        pick = np.where(rng.random(batch_size) < 0.7, kits.index(main), rng.integers(0, len(kits), size=batch_size))
        idx = np.arange(b * batch_size, (b + 1) * batch_size)
        for f in parts:
            stack = np.stack([pools[k][f][idx] for k in kits])
            parts[f].append(stack[pick, np.arange(batch_size)])
    data = {f: np.ascontiguousarray(np.concatenate(v)) for f, v in parts.items()}
    short = rng.integers(0, len(data["wlen"]), size=len(data["wlen"]) // 50)
    newlen = rng.integers(0, 150, size=len(short)).astype(np.int32)
    data["wlen"][short] = newlen
    data["read_len"][short] = newlen
    for i, n in zip(short, newlen):
        data["tail3"][i, :n] = data["win5"][i, :n]
        data["win5"][i, n:] = 0
        data["tail3"][i, n:] = 0
    return data


@pytest.fixture(scope="module")
def auto_plan():
    from qcat_b200 import config, engine, scanner
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit=None)                 # every auto-detect layout: `-k auto`
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    assert len(set(tables.kit_names)) > 4
    plan = engine.DevicePlan(tables, device=0)
    yield sc, tables, plan
    plan.close()


@pytest.mark.parametrize("force_generic", [False, True], ids=["packed", "generic"])
def test_one_pass_equals_vote_then_detect(auto_plan, force_generic):
    sc, tables, plan = auto_plan
    batch_size = 500
    data = _mixed_batches(13, batch_size, seed=5)
    n = len(data["wlen"]) - 123                                  # ragged last batch
    data = {k: v[:n] for k, v in data.items()}
    plan.set_force_generic(force_generic)
    try:
        kit_names, kit_of_layout = tables.kit_index()
        got, kits = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, batch_size,
                                     return_kits=True)
        want, want_kits = helpers.oracle_detect_auto(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"],
                                                     batch_size=batch_size, return_kits=True)
        np.testing.assert_array_equal(kits, want_kits)
        assert len(set(kits.tolist())) >= 4
        helpers.assert_records_equal(got, want, "one-pass auto kit vs oracle")
        # the two-pass route through the older entry points gives the same records
        vote = plan.kit_vote(data["win5"], data["tail3"], data["wlen"])
        for lo in range(0, n, batch_size):
            hi = min(n, lo + batch_size)
            kit = helpers.kit_from_votes(vote[lo:hi], tables.kit_names)
            part = plan.detect(data["win5"][lo:hi], data["tail3"][lo:hi], data["wlen"][lo:hi], data["read_len"][lo:hi],
                               tables.kit_subset(kit))
            helpers.assert_records_equal(got[lo:hi], part, "one-pass vs two-pass, batch at %d" % lo)
    finally:
        plan.set_force_generic(False)


def test_vote_ties_go_to_the_kit_seen_first(auto_plan):
    """Batches of two reads of different kits: one vote each, the kit of the first read wins (dict insertion order +
    stable sort, scanner_base.py:657-660)."""
    from qcat_b200 import scanner, synth
    sc, tables, plan = auto_plan
    a = synth.generate(scanner.BarcodeScannerEPI2ME(kit="PBC096").layouts, 300, seed=1, p_none=0.0, p_foreign=0.0, sub=0.01, dele=0.0, ins=0.0)
    b = synth.generate(scanner.BarcodeScannerEPI2ME(kit="RBK004").layouts, 300, seed=2, p_none=0.0, p_foreign=0.0, sub=0.01, dele=0.0, ins=0.0)
    data = {}
    for f in ("win5", "tail3", "wlen", "read_len"):
        inter = np.empty((600,) + a[f].shape[1:], dtype=a[f].dtype)
        inter[0::4], inter[1::4], inter[2::4], inter[3::4] = a[f][0::2], b[f][0::2], b[f][1::2], a[f][1::2]
        data[f] = inter
    kit_names, kit_of_layout = tables.kit_index()
    got, kits = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 2, return_kits=True)
    want, want_kits = helpers.oracle_detect_auto(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"], batch_size=2,
                                                 return_kits=True)
    np.testing.assert_array_equal(kits, want_kits)
    named = [kit_names[k] for k in kits.tolist()]
    assert named[0::2].count("PBC096") > 100 and named[1::2].count("RBK004") > 100      # the first read's kit
    helpers.assert_records_equal(got, want, "tie batches")


def test_batches_larger_than_a_device_chunk(auto_plan):
    """A batch that spans several device chunks (an API call with one huge batch) takes the vote-first route; results
    still equal per-batch vote + restricted detection.  Also checks the device-pointer entry point."""
    import torch
    from qcat_b200 import _ffi
    sc, tables, plan = auto_plan
    base = _mixed_batches(4, 2500, seed=11, kits=("RBK004", "PBC096"))
    reps = 62                                                    # 620 000 reads, batches of 300 000
    data = {k: np.ascontiguousarray(np.concatenate([v] * reps)) for k, v in base.items()}
    n, batch_size = len(data["wlen"]), 300000
    kit_names, kit_of_layout = tables.kit_index()
    got, kits = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, batch_size,
                                 return_kits=True)
    vote = plan.kit_vote(data["win5"], data["tail3"], data["wlen"])
    for b, lo in enumerate(range(0, n, batch_size)):
        hi = min(n, lo + batch_size)
        kit = helpers.kit_from_votes(vote[lo:hi], tables.kit_names)
        assert kit_names[int(kits[b])] == kit
        part = plan.detect(data["win5"][lo:hi], data["tail3"][lo:hi], data["wlen"][lo:hi], data["read_len"][lo:hi],
                           tables.kit_subset(kit))
        helpers.assert_records_equal(got[lo:hi], part, "huge batch at %d" % lo)
    head = helpers.oracle_detect(tables, data["win5"][:3000], data["tail3"][:3000], data["wlen"][:3000], data["read_len"][:3000],
                                 subset=tables.kit_subset(kit_names[int(kits[0])]))
    helpers.assert_records_equal(got[:3000], head, "huge batch head vs oracle")
    # device-resident entry point, CLI batches of 4000
    d = {k: torch.from_numpy(v).cuda() for k, v in data.items()}
    d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    d_kits = torch.zeros((n + 3999) // 4000, dtype=torch.int32, device="cuda")
    plan.detect_auto_device(d["win5"].data_ptr(), d["tail3"].data_ptr(), data["win5"].shape[1], d["wlen"].data_ptr(),
                            d["read_len"].data_ptr(), n, kit_of_layout, 4000, d_out.data_ptr(), d_batch_kit=d_kits.data_ptr(),
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host, host_kits = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 4000,
                                       return_kits=True)
    helpers.assert_records_equal(d_out.cpu().numpy().view(_ffi.RESULT_DTYPE), host, "device vs host entry point")
    np.testing.assert_array_equal(d_kits.cpu().numpy(), host_kits)
