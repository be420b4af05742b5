"""GPU: simple mode (`--simple`, scanner_simple.py) and sg_stats on the CUDA path."""
import contextlib
import io
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


def _cases():
    return [c["name"] for c in helpers.load_golden_simple()[1]]


@pytest.mark.parametrize("case_index", range(len(_cases())), ids=_cases())
def test_cuda_matches_reference_golden_simple(golden, case_index):
    from qcat_b200 import engine
    data, _, _ = golden
    sdata, cases = helpers.load_golden_simple()
    case = cases[case_index]
    idx, want = sdata["idx_%d" % case_index], sdata["res_%d" % case_index]
    tables, sc = helpers.simple_tables_for_case(case)
    plan = engine.DevicePlan(tables, device=0)
    try:
        got = plan.detect(data["win5"][idx], data["tail3"][idx], data["wlen"][idx], data["read_len"][idx])
        helpers.assert_records_equal(got, want, case["name"])
        assert plan.info()["kernel_launches"] > 0 and plan.info()["fast_barcode"] == 0
    finally:
        plan.close()


def test_simple_mode_matches_oracle_on_synthetic():
    from qcat_b200 import config, engine, scanner, synth
    from qcat_b200.tables import Tables
    for kit, src in (("standard", "NBD103/NBD104"), ("extended", "PBC096")):
        sc = scanner.BarcodeScannerSimple(kit=kit)
        data = synth.generate(scanner.BarcodeScannerEPI2ME(kit=src).layouts, 6000, seed=23)
        short = np.arange(0, 6000, 40)
        data["wlen"][short] = (short % 150).astype(np.int32)
        data["read_len"][short] = data["wlen"][short]
        for i in short:
            n = int(data["wlen"][i])
            data["tail3"][i, :n] = data["win5"][i, :n]
            data["win5"][i, n:] = 0
            data["tail3"][i, n:] = 0
        tables = Tables.simple(sc.barcodes, config.qcatConfig(), sc.min_quality)
        plan = engine.DevicePlan(tables, device=0)
        try:
            got = plan.detect(data["win5"], data["tail3"], data["wlen"], data["read_len"])
            want = helpers.oracle_detect(tables, data["win5"], data["tail3"], data["wlen"], data["read_len"])
            helpers.assert_records_equal(got, want, "simple %s" % kit)
            assert (got["barcode"] >= 0).mean() > 0.4 and (got["layout"] == -1).all()
            # stand-alone windows of any length (BarcodeScanner.scan)
            wins = [bytes(data["win5"][i, :int(data["wlen"][i])]) * (1 + i % 3) for i in range(0, 600)]
            helpers.assert_records_equal(plan.scan_windows(wins), helpers.oracle_scan(tables, wins), "simple scan %s" % kit)
            # device histogram: simple records have no layout
            import torch
            base, n_bins = plan.histogram_layout()
            d_res = torch.from_numpy(got.view(np.uint8).reshape(-1).copy()).cuda()
            d_counts = torch.zeros(n_bins, dtype=torch.int64, device="cuda")
            plan.histogram_device(d_res.data_ptr(), len(got), base, d_counts.data_ptr(), n_bins,
                                  stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            counts = d_counts.cpu().numpy()
            assert counts.sum() == len(got) and counts[0] == int((got["barcode"] < 0).sum()) and n_bins == len(sc.barcodes) + 1
        finally:
            plan.close()


def test_sg_stats_batch_matches_oracle():
    from qcat_b200 import config, engine
    cfg = config.qcatConfig()
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGTNacgtX", dtype=np.uint8)
    queries = [bytes(acgt[rng.integers(0, 10 if i % 9 == 0 else 4, size=int(rng.integers(0, 200)))]) for i in range(160)]
    refs = [bytes(acgt[rng.integers(0, 4, size=int(rng.integers(1, 110)))]) for _ in range(24)]
    for i in range(0, 160, 3):                                  # real hits: a reference planted in the query
        r = refs[i % len(refs)]
        queries[i] = queries[i][:20] + r + queries[i][20:]
    for matrix, go, ge in ((cfg.matrix_barcode, 1, 1), (cfg.matrix, cfg.gap_open, cfg.gap_extend), (cfg.matrix, 5, 1)):
        got = engine.sg_batch(queries, refs, go, ge, matrix, device=0, stats=True)
        plain = engine.sg_batch(queries, refs, go, ge, matrix, device=0)
        for a, b in zip(got[:3], plain):
            np.testing.assert_array_equal(a, b)
        for qi, q in enumerate(queries):
            for ri, r in enumerate(refs):
                want = helpers.oracle_sg_stats(q, r, go, ge, matrix)
                assert tuple(int(g[qi, ri]) for g in got) == want, (qi, ri, go, ge)


def _run_cli(argv):
    from qcat import cli
    out = io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
        cli.main(argv)
    return out.getvalue()


@pytest.mark.skipif(not refloader.available(), reason="reference package not available")
@pytest.mark.parametrize("barcodes", ["standard", "extended"])
def test_simple_mode_under_the_reference_cli(tmp_path, barcodes):
    """`qcat --simple` with the drop-in installed (BarcodeScannerSimple patched like the other scanners) and the native
    demux_file both reproduce the unmodified reference CLI: TSV and per-barcode files."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin, fastx, scanner, synth
    dropin.uninstall()
    layouts = ref_scanner.factory(kit="NBD103/NBD104" if barcodes == "standard" else "PBC096").layouts
    reads = synth.windows_to_reads(synth.generate(layouts, 500, seed=29, mean_len=800.0)) + ["", "ACGT" * 40]
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d x=%d\n%s\n+\n%s\n" % (i, i, r, "I" * len(r)))
    argv = ["-f", str(fastq), "--simple", "--simple-barcodes", barcodes, "--trim", "--tsv"]
    tsv_cpu = _run_cli(argv + ["-b", str(tmp_path / "cpu")])
    files_cpu = {name: open(tmp_path / "cpu" / name).read() for name in sorted(os.listdir(tmp_path / "cpu"))}
    dropin.install(device=0)
    try:
        tsv_gpu = _run_cli(argv + ["-b", str(tmp_path / "gpu")])
    finally:
        dropin.uninstall()
    files_gpu = {name: open(tmp_path / "gpu" / name).read() for name in sorted(os.listdir(tmp_path / "gpu"))}
    assert tsv_gpu == tsv_cpu and files_gpu == files_cpu and len(files_cpu) > 5
    sc = scanner.factory(mode="simple", kit=barcodes, device=0)
    tsv = io.StringIO()
    summary = fastx.demux_file(str(fastq), sc, trim=True, min_read_length=100, tsv=tsv, out_dir=str(tmp_path / "native"))
    files_native = {name: open(tmp_path / "native" / name).read() for name in sorted(os.listdir(tmp_path / "native"))}
    assert tsv.getvalue() == tsv_cpu and files_native == files_cpu
    assert summary["reads"] == len(reads)
