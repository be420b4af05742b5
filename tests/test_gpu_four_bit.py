"""GPU: the 4-bit window format (two base classes per byte, qcb_plan_base_classes / qcb_detect4 / qcb_detect_auto4) gives
the records of the ASCII entry points on every kind of input byte, on the packed and on the generic kernels."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def _noisy(data, seed):
    """Sprinkle lower case, IUPAC codes, N, X and junk bytes over synthetic windows."""
    rng = np.random.default_rng(seed)
    pool = np.frombuffer(b"acgtnNRYMKrymkVBHDXx*-?\x00\xff", dtype=np.uint8)
    for key in ("win5", "tail3"):
        w = data[key]
        mask = rng.random(w.shape) < 0.03
        w[mask] = pool[rng.integers(0, len(pool), size=int(mask.sum()))]
        lower = rng.random(w.shape[0]) < 0.1
        w[lower] = np.where((w[lower] >= 65) & (w[lower] <= 90), w[lower] + 32, w[lower])
    return data


@pytest.mark.parametrize("mode,kit", [("epi2me", "PBC096"), ("epi2me", "NBD103/NBD104"), ("dual", None), ("simple", "standard")])
@pytest.mark.parametrize("force_generic", [False, True], ids=["packed", "generic"])
def test_four_bit_records_equal_ascii_records(mode, kit, force_generic):
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    if mode == "simple":
        sc = scanner.BarcodeScannerSimple(kit=kit)
        tables = Tables.simple(sc.barcodes, config.qcatConfig(), sc.min_quality)
        layouts = scanner.BarcodeScannerEPI2ME(kit="NBD103/NBD104").layouts
    else:
        sc = (scanner.BarcodeScannerDual if mode == "dual" else scanner.BarcodeScannerEPI2ME)(kit=kit)
        tables = Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)
        layouts = sc.layouts
    data = _noisy(synth.generate(layouts, 70001, seed=31), seed=32)
    short = np.arange(0, 70001, 53)
    data["wlen"][short] = (short % 151).astype(np.int32)
    data["read_len"][short] = data["wlen"][short]
    plan = engine.DevicePlan(tables, device=0)
    try:
        cls = plan.base_classes()
        assert cls is not None and cls.max() < 16
        plan.set_force_generic(force_generic)
        want = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
        win5p, tail3p = plan.pack4(data["win5"], data["wlen"]), plan.pack4(data["tail3"], data["wlen"])
        assert win5p.shape[1] == 80
        got = plan.detect4(win5p, tail3p, data["wlen"], data["read_len"])
        helpers.assert_records_equal(got, want, "4-bit vs ASCII windows, %s %s" % (mode, kit))
        head = helpers.oracle_detect(tables, data["win5"][:3000], data["tail3"][:3000], data["wlen"][:3000], data["read_len"][:3000])
        helpers.assert_records_equal(got[:3000], head, "4-bit windows vs oracle")
        if mode != "simple":
            sub = plan.detect4(win5p[:5000], tail3p[:5000], data["wlen"][:5000], data["read_len"][:5000], subset=[1])
            helpers.assert_records_equal(sub, plan.detect(data["win5"][:5000], data["tail3"][:5000], data["wlen"][:5000],
                                                          data["read_len"][:5000], subset=[1]), "4-bit, layout subset")
    finally:
        plan.close()


def test_four_bit_auto_kit_and_device_entry_point():
    import torch
    from qcat_b200 import _ffi, config, engine, scanner
    from qcat_b200.tables import Tables
    from tests.test_gpu_auto_kit import _mixed_batches
    sc = scanner.BarcodeScannerEPI2ME(kit=None)
    tables = Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality)
    data = _noisy(_mixed_batches(9, 4000, seed=23, kits=("PBC096", "RBK004", "NBD103/NBD104")), seed=5)
    n = len(data["wlen"]) - 777
    data = {k: np.ascontiguousarray(v[:n]) for k, v in data.items()}
    plan = engine.DevicePlan(tables, device=0)
    try:
        kit_names, kit_of_layout = tables.kit_index()
        want, want_kits = plan.detect_auto(data["win5"], data["tail3"], data["wlen"], data["read_len"], kit_of_layout, 4000,
                                           return_kits=True)
        win5p, tail3p = plan.pack4(data["win5"], data["wlen"]), plan.pack4(data["tail3"], data["wlen"])
        got, kits = plan.detect_auto4(win5p, tail3p, data["wlen"], data["read_len"], kit_of_layout, 4000, return_kits=True)
        np.testing.assert_array_equal(kits, want_kits)
        helpers.assert_records_equal(got, want, "4-bit auto kit")
        # device-resident 4-bit windows, explicit kit subset
        subset = tables.kit_subset("PBC096")
        d = {"win5p": torch.from_numpy(win5p).cuda(), "tail3p": torch.from_numpy(tail3p).cuda(),
             "wlen": torch.from_numpy(data["wlen"]).cuda(), "read_len": torch.from_numpy(data["read_len"]).cuda()}
        d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
        plan.detect4_device(d["win5p"].data_ptr(), d["tail3p"].data_ptr(), 80, d["wlen"].data_ptr(), d["read_len"].data_ptr(), n,
                            d_out.data_ptr(), subset=subset, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        helpers.assert_records_equal(d_out.cpu().numpy().view(_ffi.RESULT_DTYPE),
                                     plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"], subset),
                                     "4-bit device entry point")
    finally:
        plan.close()
