"""GPU + reference: the UNMODIFIED qcat package (baseline/_ref or /root/reference, over the parasail stand-in) with
qcat_b200.dropin installed gives the same results as its own CPU path -- API level and through qcat.cli."""
import contextlib
import io
import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refloader.available(), reason="reference package not available")]


def _reads(kit_layouts, n, seed):
    from qcat_b200 import synth
    data = synth.generate(kit_layouts, n, seed=seed, mean_len=1200.0)
    return synth.windows_to_reads(data)


def _key(result):
    b, a = result["barcode"], result["adapter"]
    return (None if b is None else (b.name, b.id), result["barcode_score"], None if a is None else a.kit,
            result["adapter_end"], result["trim5p"], result["trim3p"], result["exit_status"])


@pytest.mark.parametrize("mode,kit,batch", [("epi2me", "PBC096", True), ("epi2me", None, True), ("epi2me", "NBD103/NBD104", False),
                                            ("dual", None, True)])
def test_dropin_matches_reference_cpu_path(mode, kit, batch):
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin
    dropin.uninstall()
    cfg = ref_config.qcatConfig()
    cpu = ref_scanner.factory(mode=mode, kit=kit)
    reads = _reads(cpu.layouts if kit or mode == "dual" else ref_scanner.factory(kit="RBK004").layouts, 240, seed=77)
    reads += ["", "ACGT", "N" * 300]
    if batch:
        want = cpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
    else:
        want = [cpu.detect_barcode(r, None, cfg) for r in reads]
    dropin.install(device=0)
    try:
        gpu = ref_scanner.factory(mode=mode, kit=kit)
        assert type(gpu) is type(cpu)
        if batch:
            got = gpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
        else:
            got = [gpu.detect_barcode(r, None, cfg) for r in reads]
        info = next(iter(gpu._qcb_plans.values())).info()
        assert info["kernel_launches"] > 0
    finally:
        dropin.uninstall()
    assert [_key(r) for r in got] == [_key(r) for r in want]
    # identity: results reference the scanner's own Barcode / AdapterLayout objects
    for r in got:
        if r["adapter"] is not None:
            assert any(r["adapter"] is l for l in gpu.layouts)
    assert "detect_barcode_batch" not in type(cpu).__dict__ or True


def _run_cli(argv):
    from qcat import cli
    out, err = io.StringIO(), io.StringIO()
    old = sys.argv
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        try:
            cli.main(argv)
        finally:
            sys.argv = old
    return out.getvalue()


def test_cli_runs_unchanged_on_top_of_the_dropin(tmp_path):
    """qcat.cli (cli.py:445-563) demultiplexes a FASTQ identically with and without the drop-in: TSV, per-barcode
    FASTQ files and trimming."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin
    dropin.uninstall()
    layouts = ref_scanner.factory(kit="PBC096").layouts
    reads = _reads(layouts, 300, seed=5)
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d truebc=x\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    outputs = {}
    for label in ("cpu", "gpu"):
        if label == "gpu":
            dropin.install(device=0)
        try:
            outdir = tmp_path / label
            tsv = _run_cli(["-f", str(fastq), "-k", "PBC096", "--tsv", "--trim", "-b", str(outdir)])
            files = {name: open(os.path.join(outdir, name)).read() for name in sorted(os.listdir(outdir))}
            outputs[label] = (tsv, files)
        finally:
            dropin.uninstall()
    assert outputs["cpu"][0] == outputs["gpu"][0]
    assert outputs["cpu"][1] == outputs["gpu"][1]
    assert len(outputs["gpu"][1]) > 10          # many per-barcode files were written


@pytest.mark.parametrize("kit,trim", [("PBC096", True), (None, False)])
def test_native_demux_file_matches_reference_cli(tmp_path, kit, trim):
    """qcat_b200.fastx.demux_file (native ingest + GPU scoring + native record writer) produces the same TSV and the
    same per-barcode FASTQ files as the unmodified reference CLI on its CPU path, batch mode included (auto kit)."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin, fastx
    dropin.uninstall()
    layouts = ref_scanner.factory(kit=kit or "RBK004").layouts
    reads = _reads(layouts, 9000 if kit is None else 600, seed=11)        # > 2 CLI batches of 4000 in auto mode
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d ch=%d\tstart=%d\n%s\n+\n%s\n" % (i, i % 512, i * 7, r, "#" * len(r)))
    argv = ["-f", str(fastq), "--tsv", "-b", str(tmp_path / "cpu"), "--min-read-length", "200"]
    if kit:
        argv += ["-k", kit]
    if trim:
        argv += ["--trim"]
    tsv_cpu = _run_cli(argv)
    files_cpu = {name: open(tmp_path / "cpu" / name).read() for name in sorted(os.listdir(tmp_path / "cpu"))}

    dropin.install(device=0)
    try:
        sc = ref_scanner.factory(kit=kit)
        tsv = io.StringIO()
        summary = fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=200, out_dir=str(tmp_path / "gpu"), tsv=tsv,
                                   chunk_bytes=1 << 20)
    finally:
        dropin.uninstall()
    files_gpu = {name: open(tmp_path / "gpu" / name).read() for name in sorted(os.listdir(tmp_path / "gpu"))}
    assert tsv.getvalue() == tsv_cpu
    assert files_gpu == files_cpu
    assert summary["reads"] == len(reads) and sum(summary["barcodes"].values()) == len(reads) - summary["skipped"]


@pytest.mark.parametrize("kit,trim,filter_barcodes", [("NBD103/NBD104", True, False), (None, False, True), ("PBC096", True, True)])
def test_native_demux_stream_output_matches_reference_cli(tmp_path, kit, trim, filter_barcodes):
    """Without -b the CLI writes every read to one stream with ' barcode=<id>' comments (cli.py:337-352); demux_file's
    `output` must be byte-identical, also with --filter-barcodes (per-batch filter, trims reset on filtered reads) and
    through the parallel record index (chunks > 1 MiB)."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin, fastx
    dropin.uninstall()
    layouts = ref_scanner.factory(kit=kit or "RBK004").layouts
    reads = _reads(layouts, 4600, seed=23)
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d%s\n%s\n+\n%s\n" % (i, " ch=%d" % (i % 512) if i % 5 else "", r, "@" * len(r)))
    argv = ["-f", str(fastq), "--min-read-length", "150"] + (["-k", kit] if kit else []) + (["--trim"] if trim else [])
    argv += ["--filter-barcodes"] if filter_barcodes else []
    stream_cpu = _run_cli(argv)
    tsv_cpu = _run_cli(argv + ["--tsv"])

    dropin.install(device=0)
    try:
        sc = ref_scanner.factory(kit=kit, enable_filter_barcodes=filter_barcodes)
        stream, tsv = io.BytesIO(), io.StringIO()
        summary = fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=150, output=stream, chunk_bytes=2 << 20)
        fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=150, tsv=tsv, chunk_bytes=3 << 20)
    finally:
        dropin.uninstall()
    assert stream.getvalue().decode() == stream_cpu
    assert tsv.getvalue() == tsv_cpu
    assert summary["reads"] == len(reads)


@pytest.mark.parametrize("mode,kit", [("epi2me", "PBC096"), ("epi2me", None), ("dual", None)])
def test_middle_adapter_scan_matches_reference(mode, kit):
    """--detect-middle (scanner_base.py:479-519, :593-595): chimeric reads (two reads joined, so adapters sit in the
    body) get exit_status 997 exactly where the reference's CPU path says so, single-read and batch API; the drop-in
    scores all bodies of a batch in one device call."""
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin
    dropin.uninstall()
    cfg = ref_config.qcatConfig()
    cpu = ref_scanner.factory(mode=mode, kit=kit, scan_middle_adapter=True)
    # low error rates: at the default ones the adapter rarely scores > 90 and the reference then only looks at
    # body[:150] for the barcode (scanner_epi2me.py:74-82), so hardly any middle adapter is found
    from qcat_b200 import synth
    layouts = cpu.layouts if kit or mode == "dual" else ref_scanner.factory(kit="RBK004").layouts
    plain = synth.windows_to_reads(synth.generate(layouts, 120, seed=31, mean_len=1200.0, sub=0.02, dele=0.01, ins=0.01))
    reads = []
    for i in range(0, len(plain), 2):
        reads.append(plain[i] + plain[i + 1] if i % 3 else plain[i])         # two thirds chimeric
    reads += ["", "ACGT" * 60, plain[0][:310], plain[1][:299]]
    want = cpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
    want_single = [cpu.detect_barcode(r, None, cfg) for r in reads[:12]]
    assert sum(r["exit_status"] == 997 for r in want) >= 5
    dropin.install(device=0)
    try:
        gpu = ref_scanner.factory(mode=mode, kit=kit, scan_middle_adapter=True)
        got = gpu.detect_barcode_batch(reads, [None] * len(reads), cfg)
        got_single = [gpu.detect_barcode(r, None, cfg) for r in reads[:12]]
    finally:
        dropin.uninstall()
    assert [_key(r) for r in got] == [_key(r) for r in want]
    assert [_key(r) for r in got_single] == [_key(r) for r in want_single]


def test_module_entry_point_runs_the_reference_cli(tmp_path):
    """`python -m qcat_b200 <qcat arguments>`: the reference's own cli.main on top of the drop-in, as a subprocess (the
    parasail stand-in and the installed reference are put on PYTHONPATH the way a real parasail + qcat would be)."""
    import subprocess
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin
    dropin.uninstall()
    reads = _reads(ref_scanner.factory(kit="NBD103/NBD104").layouts, 200, seed=3)
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    want = _run_cli(["-f", str(fastq), "-k", "NBD103/NBD104", "--tsv"])
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "oracle", "refshim"), refloader.REFERENCE_ROOT,
                                         env.get("PYTHONPATH", "")])
    for extra in ([], ["--devices", "0,0"]):
        proc = subprocess.run([sys.executable, "-m", "qcat_b200"] + extra + ["-f", str(fastq), "-k", "NBD103/NBD104", "--tsv"],
                              cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
        assert proc.returncode == 0, proc.stderr[-2000:]
        assert proc.stdout == want


@pytest.mark.parametrize("options", [["--no-batch"], ["--dual"], ["--min-score", "75", "-k", "NBD103/NBD104"],
                                     ["--detect-middle", "-k", "RBK004"], ["--no-batch", "--dual", "--trim"]])
def test_cli_option_matrix_on_top_of_the_dropin(tmp_path, options):
    """The CLI switches that change what the scanner is asked (cli.py:445-563: --no-batch -> per-read detect_barcode,
    --dual, --min-score, --detect-middle): same TSV and same output stream with and without the drop-in."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import dropin
    dropin.uninstall()
    kit = options[options.index("-k") + 1] if "-k" in options else "RBK004"
    layouts = ref_scanner.factory(mode="dual" if "--dual" in options else "epi2me", kit=None if "--dual" in options else kit).layouts
    reads = _reads(layouts, 260, seed=41)
    fastq = tmp_path / "reads.fastq"
    with open(fastq, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d\n%s\n+\n%s\n" % (i, r, "5" * len(r)))
    argv = ["-f", str(fastq), "--min-read-length", "120"] + options
    want = (_run_cli(argv + ["--tsv"]), _run_cli(argv))
    dropin.install(device=0)
    try:
        got = (_run_cli(argv + ["--tsv"]), _run_cli(argv))
    finally:
        dropin.uninstall()
    assert got == want
    assert want[0].count("\n") > 200
