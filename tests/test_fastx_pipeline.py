"""CPU: qcat_b200.fastx.demux_file (native reader -> window packing -> batch logic -> native writers, three pipeline
threads) against the UNMODIFIED reference CLI, with the CPU oracle standing in for the device plan.  The oracle is the
checker's scorer here (tests only); the GPU twin of this test is tests/test_gpu_dropin.py."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from tests import helpers

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

pytestmark = pytest.mark.skipif(not refloader.available(), reason="reference package not available")


OraclePlan = helpers.OraclePlan


def _oracle_scanner(mode, kit, **kw):
    from qcat_b200 import config, scanner
    from qcat_b200.tables import Tables
    cls = scanner.BarcodeScannerDual if mode == "dual" else scanner.BarcodeScannerEPI2ME
    sc = cls(kit=kit, **kw)
    plan = OraclePlan(Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality))
    sc._plan_for = lambda qcat_config, layouts=None: plan
    return sc


def _run_cli(argv):
    from qcat import cli
    out = io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
        cli.main(argv)
    return out.getvalue()


def _write_fastq(path, reads, qual_char="#"):
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("@read%d%s\n%s\n+\n%s\n" % (i, " ch=%d\tstart=%d" % (i % 512, i * 7) if i % 4 else "", r, qual_char * len(r)))


@pytest.mark.parametrize("mode,kit,trim,filter_barcodes,n_reads", [
    ("epi2me", "PBC096", True, False, 700),
    ("epi2me", None, False, True, 4300),          # kit auto: per-batch vote over 12 layouts, two CLI batches
    ("dual", None, True, False, 300),
])
def test_demux_file_matches_reference_cli(tmp_path, mode, kit, trim, filter_barcodes, n_reads):
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import fastx, synth
    layouts = ref_scanner.factory(mode=mode, kit=kit or "RBK004").layouts if (kit or mode == "dual") else \
        ref_scanner.factory(kit="RBK004").layouts
    reads = synth.windows_to_reads(synth.generate(layouts, n_reads, seed=41, mean_len=700.0))
    reads += ["", "ACGTN" * 70]
    fastq = tmp_path / "reads.fastq"
    _write_fastq(fastq, reads)
    argv = ["-f", str(fastq), "--min-read-length", "120"] + (["-k", kit] if kit else []) + (["--trim"] if trim else [])
    argv += (["--filter-barcodes"] if filter_barcodes else []) + (["--dual"] if mode == "dual" else [])
    stream_cpu = _run_cli(argv)
    tsv_cpu = _run_cli(argv + ["--tsv", "-b", str(tmp_path / "cpu")])
    files_cpu = {name: open(tmp_path / "cpu" / name).read() for name in sorted(os.listdir(tmp_path / "cpu"))}

    sc = _oracle_scanner(mode, kit, enable_filter_barcodes=filter_barcodes)
    stream, tsv = io.BytesIO(), io.StringIO()
    first = fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=120, output=stream, chunk_bytes=1 << 20)
    second = fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=120, tsv=tsv, out_dir=str(tmp_path / "native"),
                              chunk_bytes=300000)
    files_native = {name: open(tmp_path / "native" / name).read() for name in sorted(os.listdir(tmp_path / "native"))}
    assert stream.getvalue().decode() == stream_cpu
    assert tsv.getvalue() == tsv_cpu
    assert files_native == files_cpu
    assert first["reads"] == second["reads"] == len(reads)
    # `-o file`: output given as a path
    fastx.demux_file(str(fastq), sc, trim=trim, min_read_length=120, output=str(tmp_path / "stream.out"), chunk_bytes=150000)
    assert (tmp_path / "stream.out").read_text() == stream_cpu
    assert first["barcodes"] == second["barcodes"] and sum(first["barcodes"].values()) == len(reads) - first["skipped"]
    np.testing.assert_array_equal(first["records"], second["records"])


@pytest.mark.parametrize("fmt", ["fasta_wrapped", "fastq_crlf"])
def test_demux_file_other_input_layouts(tmp_path, fmt):
    """Wrapped FASTA (the usual 60-column layout) and CRLF line ends go through the same native reader, window packing
    and writers; output must equal the reference CLI's (Bio's parsers join the rstrip()ped lines).  Wrapped FASTQ is
    covered natively in tests/test_fastx.py only: the Bio stand-in of oracle/refshim reads four-line records."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import fastx, synth
    layouts = ref_scanner.factory(kit="PBC096").layouts
    reads = synth.windows_to_reads(synth.generate(layouts, 400, seed=43, mean_len=500.0)) + ["ACGT" * 20, "A"]
    path = tmp_path / ("reads.fasta" if fmt.startswith("fasta") else "reads.fastq")
    with open(path, "w", newline="") as fh:
        for i, r in enumerate(reads):
            if fmt == "fasta_wrapped":
                fh.write(">read%d desc=%d\n" % (i, i) + "".join(r[j:j + 60] + "\n" for j in range(0, len(r), 60)))
            else:
                fh.write("@read%d x\r\n%s\r\n+\r\n%s\r\n" % (i, r, "5" * len(r)))
    argv = ["-f", str(path), "-k", "PBC096", "--trim", "--min-read-length", "50"]
    stream_cpu = _run_cli(argv)
    tsv_cpu = _run_cli(argv + ["--tsv", "-b", str(tmp_path / "cpu")])
    files_cpu = {name: open(tmp_path / "cpu" / name).read() for name in sorted(os.listdir(tmp_path / "cpu"))}
    sc = _oracle_scanner("epi2me", "PBC096")
    stream, tsv = io.BytesIO(), io.StringIO()
    fastx.demux_file(str(path), sc, trim=True, min_read_length=50, output=stream, chunk_bytes=70000)
    fastx.demux_file(str(path), sc, trim=True, min_read_length=50, tsv=tsv, out_dir=str(tmp_path / "native"), chunk_bytes=1 << 20)
    files_native = {name: open(tmp_path / "native" / name).read() for name in sorted(os.listdir(tmp_path / "native"))}
    assert stream.getvalue().decode() == stream_cpu
    assert tsv.getvalue() == tsv_cpu
    assert files_native == files_cpu and len(files_native) > 5


def test_demux_file_propagates_errors(tmp_path):
    """A malformed record deep inside the file and a failing scorer both surface as exceptions on the caller's thread
    (no hang in the pipeline threads)."""
    from qcat_b200 import fastx, synth
    sc = _oracle_scanner("epi2me", "PBC096")
    reads = synth.windows_to_reads(synth.generate(sc.layouts, 600, seed=2, mean_len=600.0))
    good = tmp_path / "good.fastq"
    _write_fastq(good, reads)
    bad = tmp_path / "bad.fastq"
    text = good.read_text().split("\n")
    text[4 * 400 + 3] = text[4 * 400 + 3][:-3]                      # quality shorter than the sequence in record 400
    bad.write_text("\n".join(text))
    with pytest.raises(fastx.FastxError):
        fastx.demux_file(str(bad), sc, chunk_bytes=100000)

    class Boom(RuntimeError):
        pass

    def explode(*args, **kwargs):
        raise Boom("scorer failed")
    sc._plan_for(None).detect = explode
    with pytest.raises(Boom):
        fastx.demux_file(str(good), sc, chunk_bytes=100000)
    with pytest.raises(IOError):
        fastx.demux_file(str(tmp_path / "missing.fastq"), sc)


@pytest.mark.parametrize("kit,filter_barcodes", [("PBC096", False), (None, True)])
def test_demux_file_detect_middle_matches_reference_cli(tmp_path, kit, filter_barcodes):
    """`--detect-middle` through demux_file (scanner_base.py:479-519, :593-595): chimeric reads become `none` (exit
    status 997) exactly where the reference CLI says so -- before the per-batch barcode filter, wrapped FASTA included."""
    refloader.load()
    from qcat import scanner as ref_scanner
    from qcat_b200 import fastx, synth
    layouts = ref_scanner.factory(kit=kit or "RBK004").layouts
    plain = synth.windows_to_reads(synth.generate(layouts, 160, seed=37, mean_len=900.0, sub=0.02, dele=0.01, ins=0.01))
    reads = [plain[i] + plain[i + 1] if i % 3 else plain[i] for i in range(0, len(plain), 2)]
    reads += ["", "ACGT" * 60, plain[0][:310], plain[1][:299]]
    path = tmp_path / "reads.fasta"
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            fh.write(">read%d\n" % i + ("".join(r[j:j + 70] + "\n" for j in range(0, len(r), 70)) or "\n"))
    argv = ["-f", str(path), "--detect-middle"] + (["-k", kit] if kit else []) + (["--filter-barcodes"] if filter_barcodes else [])
    tsv_cpu = _run_cli(argv + ["--tsv", "-b", str(tmp_path / "cpu")])
    files_cpu = {name: open(tmp_path / "cpu" / name).read() for name in sorted(os.listdir(tmp_path / "cpu"))}
    sc = _oracle_scanner("epi2me", kit, enable_filter_barcodes=filter_barcodes, scan_middle_adapter=True)
    tsv = io.StringIO()
    summary = fastx.demux_file(str(path), sc, tsv=tsv, out_dir=str(tmp_path / "native"), chunk_bytes=60000,
                               min_read_length=100)              # the CLI's default
    files_native = {name: open(tmp_path / "native" / name).read() for name in sorted(os.listdir(tmp_path / "native"))}
    assert tsv.getvalue() == tsv_cpu
    assert files_native == files_cpu
    assert int((summary["records"]["exit_status"] == 997).sum()) >= 5
    # without the flag the same reads stay classified: the scan is what made the difference
    plain_sc = _oracle_scanner("epi2me", kit, enable_filter_barcodes=filter_barcodes)
    assert int((fastx.demux_file(str(path), plain_sc)["records"]["exit_status"] == 997).sum()) == 0
