"""GPU: several devices behind the API inside one process (MultiDevicePlan, qcb_detect_multi / qcb_detect_auto_multi,
qcb_hist_allgather).  With one visible GPU the two plans share device 0 -- the sharding, threading and write-back logic
is the same; with two or more (gpurun --gpus 2) they sit on different devices."""
import os
import sys

import numpy as np
import pytest

from tests import helpers

ROOT = helpers.ROOT
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

pytestmark = pytest.mark.gpu


def _devices():
    from qcat_b200 import engine
    n = engine.device_count()
    assert n > 0
    return list(range(min(n, 4))) if n >= 2 else [0, 0]


def _ragged(data, every=37):
    idx = np.arange(0, len(data["wlen"]), every)
    data["wlen"][idx] = (idx % 150).astype(np.int32)
    data["read_len"][idx] = data["wlen"][idx]
    for i in idx:
        n = int(data["wlen"][i])
        data["tail3"][i, :n] = data["win5"][i, :n]
        data["win5"][i, n:] = 0
        data["tail3"][i, n:] = 0
    return data


def test_multi_device_detect_equals_single_device():
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096")
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    data = _ragged(synth.generate(sc.layouts, 400003, seed=3))
    single = engine.DevicePlan(tables, device=0)
    multi = engine.make_plan(tables, device=_devices())
    try:
        assert isinstance(multi, engine.MultiDevicePlan)
        want = single.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        got = multi.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        helpers.assert_records_equal(got, want, "multi-device vs single-device records")
        assert all(l > 0 for l in multi.info()["kernel_launches_per_device"]), "a device got no work"
        head = helpers.oracle_detect(tables, data["win5"][:3000], data["tail3"][:3000], data["wlen"][:3000], data["read_len"][:3000])
        helpers.assert_records_equal(got[:3000], head, "multi-device vs oracle")
        small = multi.detect(data["win5"][:1000], data["tail3"][:1000], data["wlen"][:1000], data["read_len"][:1000], subset=[1])
        helpers.assert_records_equal(small, single.detect(data["win5"][:1000], data["tail3"][:1000], data["wlen"][:1000],
                                                          data["read_len"][:1000], subset=[1]), "small call, layout subset")

        # the path's one exchange: per-device count vectors all-gathered with peer copies (qcb_hist_allgather)
        import torch
        base, n_bins = multi.histogram_layout()
        devs = multi.devices
        shards = np.array_split(np.arange(len(got)), len(devs))
        counts, gathered = [], []
        for d, idx in zip(devs, shards):
            with torch.cuda.device(d):
                res = torch.from_numpy(got[idx].view(np.uint8).reshape(-1).copy()).cuda(d)
                c = torch.zeros(n_bins, dtype=torch.int64, device="cuda:%d" % d)
                multi.plans[len(counts)].histogram_device(res.data_ptr(), len(idx), base, c.data_ptr(), n_bins,
                                                          stream=torch.cuda.current_stream(d).cuda_stream)
                torch.cuda.synchronize(d)
                counts.append(c)
                gathered.append(torch.zeros(len(devs) * n_bins, dtype=torch.int64, device="cuda:%d" % d))
        multi.hist_allgather([c.data_ptr() for c in counts], n_bins, [g.data_ptr() for g in gathered])
        from qcat_b200 import dist as qdist
        total = qdist.histogram_bins(got, base, n_bins)
        for g in gathered:
            stack = g.cpu().numpy().reshape(len(devs), n_bins)
            np.testing.assert_array_equal(stack.sum(0), total)
            for j, idx in enumerate(shards):
                np.testing.assert_array_equal(stack[j], qdist.histogram_bins(got[idx], base, n_bins))
    finally:
        single.close()
        multi.close()


def test_multi_device_auto_kit_equals_single_device():
    from qcat_b200 import config, engine, scanner
    from qcat_b200.tables import Tables
    from tests.test_gpu_auto_kit import _mixed_batches
    sc = scanner.BarcodeScannerEPI2ME(kit=None)
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    base = _mixed_batches(6, 4000, seed=17, kits=("PBC096", "RBK004", "NBD103/NBD104"))
    data = {k: np.ascontiguousarray(np.concatenate([v] * 9)[:-1234]) for k, v in base.items()}
    kit_names, kit_of_layout = tables.kit_index()
    single = engine.DevicePlan(tables, device=0)
    multi = engine.MultiDevicePlan(tables, _devices())
    try:
        want, want_kits = single.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 4000,
                                             return_kits=True)
        got, kits = multi.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 4000,
                                      return_kits=True)
        np.testing.assert_array_equal(kits, want_kits)
        assert len(set(kits.tolist())) == 3
        helpers.assert_records_equal(got, want, "multi-device auto kit")
        assert all(l > 0 for l in multi.info()["kernel_launches_per_device"])
    finally:
        single.close()
        multi.close()


@pytest.mark.skipif(not refloader.available(), reason="reference package not available")
def test_dropin_spreads_batches_over_devices():
    """dropin.install(devices=...) under the unmodified reference classes: same results as the reference's CPU path."""
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin, engine, synth
    dropin.uninstall()
    cfg = ref_config.qcatConfig()
    cpu = ref_scanner.factory(mode="epi2me", kit="NBD103/NBD104")
    reads = synth.windows_to_reads(synth.generate(cpu.layouts, 700, seed=13)) + ["", "ACGT"]
    want = cpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
    dropin.install(devices=_devices())
    try:
        gpu = ref_scanner.factory(mode="epi2me", kit="NBD103/NBD104")
        got = gpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
        plan = next(iter(gpu._qcb_plans.values()))
        assert isinstance(plan, engine.MultiDevicePlan)
    finally:
        dropin.uninstall()

    def key(r):
        b, a = r["barcode"], r["adapter"]
        return (None if b is None else b.name, r["barcode_score"], None if a is None else a.kit, r["adapter_end"], r["trim5p"],
                r["trim3p"], r["exit_status"])
    assert [key(r) for r in got] == [key(r) for r in want]
