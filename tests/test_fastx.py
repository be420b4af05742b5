"""Native FASTQ / FASTA ingest (qcb_fastx_index, qcb_pack_windows, qcb_format_records) against plain Python."""
import io
import os
import random

import numpy as np
import pytest

from qcat_b200 import fastx
from qcat_b200.tables import pack_windows as py_pack_windows


def _random_reads(n, seed, max_len=900):
    rng = random.Random(seed)
    reads = []
    for i in range(n):
        length = rng.choice([0, 1, 5, 149, 150, 151, 299, 300]) if i % 7 == 0 else rng.randint(1, max_len)
        seq = "".join(rng.choice("ACGTN") for _ in range(length))
        qual = "".join(chr(33 + rng.randint(0, 40)) for _ in range(length))
        title = "read%d" % i + (" runid=%d\tch=%d" % (i * 3, i % 512) if i % 3 else "")
        reads.append((title, seq, qual))
    return reads


def _fastq_bytes(reads, crlf=False):
    nl = "\r\n" if crlf else "\n"
    return "".join("@%s%s%s%s+%s%s%s" % (t, nl, s, nl, nl, q, nl) for t, s, q in reads).encode()


def test_fastq_index_and_windows_match_python():
    reads = _random_reads(500, 1)
    for crlf in (False, True):
        buf = _fastq_bytes(reads, crlf)
        recs, consumed, is_fastq = fastx.index_buffer(buf)
        assert is_fastq and consumed == len(buf) and len(recs) == len(reads)
        for r, (title, seq, qual) in zip(recs, reads):
            assert buf[r["title_off"]:r["title_off"] + r["title_len"]].decode() == title
            assert r["seq_len"] == len(seq)
            assert buf[r["seq_off"]:r["seq_off"] + r["seq_span"]].decode().rstrip("\r") == seq
            assert buf[r["qual_off"]:r["qual_off"] + r["qual_span"]].decode().rstrip("\r") == qual
        got = fastx.pack_windows(buf, recs, 150)
        want = py_pack_windows([s for _, s, _ in reads], 150)[:4]
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)


def test_multiline_fasta_and_fastq():
    rng = random.Random(5)
    seqs = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (0, 1, 59, 60, 61, 150, 400, 1000)]
    fasta = "".join(">s%d desc\n%s" % (i, "".join(s[j:j + 60] + "\n" for j in range(0, len(s), 60)) or "\n") for i, s in enumerate(seqs)).encode()
    recs, consumed, is_fastq = fastx.index_buffer(fasta)
    assert not is_fastq and len(recs) == len(seqs) and [int(r["seq_len"]) for r in recs] == [len(s) for s in seqs]
    got = fastx.pack_windows(fasta, recs, 150)
    want = py_pack_windows(seqs, 150)[:4]
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    fastq = "".join("@s%d\n%s+\n%s" % (i, "".join(s[j:j + 70] + "\n" for j in range(0, len(s), 70)) or "\n",
                                         "".join("I" * len(s[j:j + 70]) + "\n" for j in range(0, len(s), 70)) or "\n")
                    for i, s in enumerate(seqs)).encode()
    recs, consumed, is_fastq = fastx.index_buffer(fastq)
    assert is_fastq and [int(r["seq_len"]) for r in recs] == [len(s) for s in seqs]
    got = fastx.pack_windows(fastq, recs, 150)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_chunked_reading_is_seamless(tmp_path):
    reads = _random_reads(3000, 2)
    path = tmp_path / "r.fastq"
    path.write_bytes(_fastq_bytes(reads)[:-1])                    # no trailing newline on the last record
    titles = []
    for buf, recs, is_fastq in fastx.iter_chunks(str(path), chunk_bytes=50000, multiple_of=128):
        titles += [buf[r["title_off"]:r["title_off"] + r["title_len"]].decode() for r in recs]
    assert titles == [t for t, _, _ in reads]
    sizes = [len(recs) for _, recs, _ in fastx.iter_chunks(str(path), chunk_bytes=50000, multiple_of=128)]
    assert all(s % 128 == 0 for s in sizes[:-1]) and sum(sizes) == 3000


def test_malformed_input_raises_like_the_reference():
    with pytest.raises(fastx.FastxError, match="must start with"):
        fastx.index_buffer(b"ACGT\n")
    with pytest.raises(fastx.FastxError, match="should start with '@'"):
        fastx.index_buffer(b"@a\nACGT\n+\nIIII\nACGT\n")
    with pytest.raises(fastx.FastxError, match="quality"):
        fastx.index_buffer(b"@a\nACGT\n+\nII\n")
    with pytest.raises(fastx.FastxError, match="Missing"):
        fastx.index_buffer(b"@a\nACGT\n")
    recs, consumed, _ = fastx.index_buffer(b"")
    assert len(recs) == 0


def test_format_records_matches_cli_layout():
    """Per-barcode record text == cli.py:write_to_file's print(...) including trimming and the min-length filter."""
    import ctypes
    from qcat_b200 import _ffi
    reads = _random_reads(300, 3)
    buf = _fastq_bytes(reads)
    recs, _, _ = fastx.index_buffer(buf)
    n = len(recs)
    rng = np.random.default_rng(0)
    results = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    lens = np.array([len(s) for _, s, _ in reads])
    results["trim5p"] = rng.integers(0, 80, size=n)
    results["trim3p"] = np.maximum(lens - rng.integers(0, 80, size=n), 0)
    bins = rng.integers(0, 4, size=n).astype(np.int32)
    lib = _ffi.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    for trim, min_len in ((0, 0), (1, 0), (1, 100)):
        bin_bytes = np.zeros(4, dtype=np.int64); bin_off = np.zeros(4, dtype=np.int64); kept = np.zeros(n, dtype=np.uint8)
        args = (vp(arr), vp(recs), vp(results), vp(bins), n, 4, 1, trim, min_len, vp(bin_bytes))
        assert lib.qcb_format_records(*args, None, 0, vp(bin_off), vp(kept), 4) == 0
        out = np.zeros(int(bin_bytes.sum()) + 1, dtype=np.uint8)
        assert lib.qcb_format_records(*args, vp(out), int(out.size), vp(bin_off), vp(kept), 4) == 0
        for b in range(4):
            want = io.StringIO()
            for i, (title, seq, qual) in enumerate(reads):
                if bins[i] != b:
                    continue
                if trim:
                    seq, qual = seq[results["trim5p"][i]:results["trim3p"][i]], qual[results["trim5p"][i]:results["trim3p"][i]]
                if len(seq) < min_len:
                    assert not kept[i]
                    continue
                cols = title.replace("\t", " ").split(" ")
                name, comment = cols[0], (" ".join(cols[1:]) if len(cols) > 1 else None)
                print("@" + name + " " + (comment or ""), seq, "+", qual or "", sep="\n", file=want)
            got = out[bin_off[b]:bin_off[b] + bin_bytes[b]].tobytes().decode()
            assert got == want.getvalue()


def _big_reads(n, seed, qual_alphabet):
    rng = random.Random(seed)
    reads = []
    for i in range(n):
        length = rng.randint(0, 3) if i % 97 == 0 else rng.randint(200, 2400)
        seq = "".join(rng.choices("ACGT", k=length))
        qual = "".join(rng.choices(qual_alphabet, k=length))
        reads.append(("r%d ch=%d" % (i, i % 512), seq, qual))
    return reads


def test_parallel_index_equals_serial_scan():
    """qcb_fastx_index_mt cuts the buffer at record starts it can verify; quality lines that start with '@' or '+'
    (the classic FASTQ ambiguity) must not fool it, and wrapped FASTQ must fall back to the serial scan."""
    reads = _big_reads(2500, 9, "@+I5#")                              # ~3.3 MB, every 5th quality line starts with '@' / '+'
    buf = _fastq_bytes(reads)
    assert len(buf) > (2 << 20)
    serial, consumed_s, _ = fastx.index_buffer(buf, threads=1)
    for threads in (2, 3, 8):
        par, consumed_p, is_fastq = fastx.index_buffer(buf, threads=threads)
        assert is_fastq and consumed_p == consumed_s == len(buf)
        assert len(par) == len(reads) and par.tobytes() == serial.tobytes()
    # not the end of the input: the incomplete last record stays unconsumed in both scans
    cutoff = len(buf) - 700
    serial, consumed_s, _ = fastx.index_buffer(buf[:cutoff], final_chunk=False, threads=1)
    par, consumed_p, _ = fastx.index_buffer(buf[:cutoff], final_chunk=False, threads=8)
    assert consumed_p == consumed_s < cutoff and par.tobytes() == serial.tobytes()
    # max_records smaller than the buffer holds
    serial, consumed_s, _ = fastx.index_buffer(buf, max_records=1000, threads=1)
    par, consumed_p, _ = fastx.index_buffer(buf, max_records=1000, threads=8)
    assert len(par) == 1000 and consumed_p == consumed_s and par.tobytes() == serial.tobytes()
    # wrapped (multi-line) FASTQ and FASTA
    wrapped = "".join("@%s\n%s+\n%s" % (t, "".join(s[j:j + 80] + "\n" for j in range(0, len(s), 80)) or "\n",
                                          "".join(q[j:j + 80] + "\n" for j in range(0, len(q), 80)) or "\n")
                      for t, s, q in reads).encode()
    serial, _, _ = fastx.index_buffer(wrapped, threads=1)
    par, _, _ = fastx.index_buffer(wrapped, threads=8)
    assert [int(x) for x in serial["seq_len"]] == [len(s) for _, s, _ in reads] and par.tobytes() == serial.tobytes()
    fasta = "".join(">%s\n%s" % (t, "".join(s[j:j + 60] + "\n" for j in range(0, len(s), 60)) or "\n") for t, s, _ in reads).encode()
    serial, _, is_fastq = fastx.index_buffer(fasta, threads=1)
    par, _, _ = fastx.index_buffer(fasta, threads=8)
    assert not is_fastq and len(par) == len(reads) and par.tobytes() == serial.tobytes()
    # errors surface identically
    broken = buf[:len(buf) // 2] + b"@bad\nACGT\n+\nII\n" + buf[len(buf) // 2:]
    with pytest.raises(fastx.FastxError):
        fastx.index_buffer(broken, threads=1)
    with pytest.raises(fastx.FastxError):
        fastx.index_buffer(broken, threads=8)


@pytest.mark.parametrize("chunk_bytes,multiple_of", [(1 << 20, 1), (1 << 20, 128), (300000, 1000), (64 << 20, 4000)])
def test_native_reader_chunks(tmp_path, chunk_bytes, multiple_of):
    reads = _big_reads(4100, 4, "I5#@")
    path = tmp_path / "r.fastq"
    path.write_bytes(_fastq_bytes(reads)[:-1])                        # last record without a line break
    seen = []
    sizes = []
    with fastx.Reader(str(path), chunk_bytes, threads=4) as reader:
        for chunk in reader.chunks(multiple_of):
            data = chunk.data.tobytes()
            assert chunk.fastq
            for r in chunk.recs:
                seen.append((data[r["title_off"]:r["title_off"] + r["title_len"]].decode(),
                             data[r["seq_off"]:r["seq_off"] + r["seq_span"]].decode(),
                             data[r["qual_off"]:r["qual_off"] + r["qual_span"]].decode()))
            sizes.append(len(chunk))
            chunk.release()
    assert seen == reads
    assert all(s % multiple_of == 0 for s in sizes[:-1]) and sum(sizes) == len(reads)


def test_native_reader_errors(tmp_path):
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(b"ACGT\n")
    with fastx.Reader(str(bad)) as reader, pytest.raises(fastx.FastxError, match="must start with"):
        reader.next_chunk()
    bad.write_bytes(b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nIIII\n")
    with fastx.Reader(str(bad)) as reader, pytest.raises(fastx.FastxError, match="quality"):
        reader.next_chunk()
    bad.write_bytes(b"@a\nACGT\n+\nIIII\nxyz\n")
    with fastx.Reader(str(bad)) as reader, pytest.raises(fastx.FastxError):
        reader.next_chunk()
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    with fastx.Reader(str(empty)) as reader:
        assert reader.next_chunk() is None
    with pytest.raises(IOError):
        fastx.Reader(str(tmp_path / "missing.fastq"))


def _py_trim(seq, qual, res, trim):
    if trim:
        return seq[res["trim5p"]:res["trim3p"]], qual[res["trim5p"]:res["trim3p"]]
    return seq, qual


def test_stream_and_tsv_writers_match_cli_layout():
    """qcb_format_stream == cli.py:write_to_file without -b (comment + ' barcode=<id>'), qcb_format_tsv ==
    cli.py:write_multiplexing_result, including repr() of the score, trimming and the min-length filter."""
    import ctypes
    from qcat_b200 import _ffi
    reads = _random_reads(400, 8)
    buf = _fastq_bytes(reads)
    recs, _, _ = fastx.index_buffer(buf)
    n = len(recs)
    rng = np.random.default_rng(2)
    results = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    lens = np.array([len(s) for _, s, _ in reads])
    results["trim5p"] = rng.integers(0, 80, size=n)
    results["trim3p"] = np.maximum(lens - rng.integers(0, 80, size=n), 0)
    results["adapter_end"] = rng.integers(0, 150, size=n)
    scores = [100.0, 0.0, 97.61904761904762, 58.0, 84.78260869565217, 2.380952380952381, 83.33333333333333, -2.1739130434782608,
              61.904761904761905, 1e-3, 123456.789]
    results["barcode_score"] = [scores[i % len(scores)] * (1.0 if i % 3 else 100.0 / 42 / 2.380952380952381) for i in range(n)]
    strings = ["none", "7", "12", "6/95", "PBC096", "NBD103/NBD104"]
    blob = np.frombuffer("".join(strings).encode() + b"\0", dtype=np.uint8)
    off = np.zeros(len(strings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in strings])
    id_label = rng.integers(0, 4, size=n).astype(np.int32)
    kit_label = rng.integers(4, 6, size=n).astype(np.int32)
    lib = _ffi.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    for trim, min_len in ((0, 0), (1, 0), (1, 100)):
        kept = np.zeros(n, dtype=np.uint8)
        need = ctypes.c_int64(0)
        args = (vp(arr), vp(recs), vp(results), vp(id_label), n, vp(blob), vp(off), len(strings), 1, trim, min_len)
        assert lib.qcb_format_stream(*args, None, 0, ctypes.byref(need), vp(kept), 4) == 0
        out = np.zeros(need.value + 1, dtype=np.uint8)
        assert lib.qcb_format_stream(*args, vp(out), int(out.size), ctypes.byref(need), vp(kept), 4) == 0
        want, want_tsv = io.StringIO(), io.StringIO()
        for i, (title, seq, qual) in enumerate(reads):
            seq, qual = _py_trim(seq, qual, results[i], trim)
            if len(seq) < min_len:
                assert not kept[i]
                continue
            cols = title.replace("\t", " ").split(" ")
            name, comment = cols[0], (" ".join(cols[1:]) if len(cols) > 1 else None)
            bc = strings[id_label[i]]
            print("@" + name + " " + "{} barcode={}".format(comment or "", bc), seq, "+", qual or "", sep="\n", file=want)
            if id_label[i]:
                print(name, len(seq), bc, float(results["barcode_score"][i]), strings[kit_label[i]],
                      int(results["adapter_end"][i]), comment, sep="\t", file=want_tsv)
            else:
                print(name, len(seq), "none", "-1", "none", "-1", comment, sep="\t", file=want_tsv)
        assert out[:need.value].tobytes().decode() == want.getvalue()
        tsv_label = np.where(id_label > 0, id_label, -1).astype(np.int32)
        cap = 200 * n
        out = np.zeros(cap, dtype=np.uint8)
        assert lib.qcb_format_tsv(vp(arr), vp(recs), vp(results), vp(tsv_label), vp(kit_label), n, vp(blob), vp(off), len(strings),
                                  trim, min_len, vp(out), cap, ctypes.byref(need), vp(kept), 4) == 0
        assert out[:need.value].tobytes().decode() == want_tsv.getvalue()


def test_writers_handle_wrapped_records():
    """qcb_format_records / qcb_format_stream on wrapped FASTQ (80 columns) and CRLF input cut [trim5p:trim3p] in base
    coordinates, line breaks skipped."""
    import ctypes
    from qcat_b200 import _ffi
    reads = _random_reads(120, 12, max_len=700)
    for layout in ("wrapped", "crlf"):
        if layout == "wrapped":
            wrap = lambda t: "".join(t[j:j + 80] + "\n" for j in range(0, len(t), 80)) or "\n"
            buf = "".join("@%s\n%s+\n%s" % (t, wrap(s), wrap(q)) for t, s, q in reads).encode()
        else:
            buf = _fastq_bytes(reads, crlf=True)
        recs, _, _ = fastx.index_buffer(buf)
        n = len(recs)
        assert n == len(reads)
        rng = np.random.default_rng(4)
        results = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        lens = np.array([len(s) for _, s, _ in reads])
        results["trim5p"] = rng.integers(0, 120, size=n)
        results["trim3p"] = np.maximum(lens - rng.integers(0, 120, size=n), 0)
        strings = ["none", "3"]
        blob = np.frombuffer(b"none3\0", dtype=np.uint8)
        off = np.array([0, 4, 5], dtype=np.int64)
        label = rng.integers(0, 2, size=n).astype(np.int32)
        lib = _ffi.load()
        arr = np.frombuffer(buf, dtype=np.uint8)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        kept = np.zeros(n, dtype=np.uint8)
        need = ctypes.c_int64(0)
        args = (vp(arr), vp(recs), vp(results), vp(label), n, vp(blob), vp(off), 2, 1, 1, 30)
        assert lib.qcb_format_stream(*args, None, 0, ctypes.byref(need), vp(kept), 3) == 0
        out = np.zeros(need.value + 1, dtype=np.uint8)
        assert lib.qcb_format_stream(*args, vp(out), int(out.size), ctypes.byref(need), vp(kept), 3) == 0
        want = io.StringIO()
        for i, (title, seq, qual) in enumerate(reads):
            seq, qual = seq[results["trim5p"][i]:results["trim3p"][i]], qual[results["trim5p"][i]:results["trim3p"][i]]
            if len(seq) < 30:
                continue
            cols = title.replace("\t", " ").split(" ")
            name, comment = cols[0], (" ".join(cols[1:]) if len(cols) > 1 else None)
            print("@" + name + " " + "{} barcode={}".format(comment or "", strings[label[i]]), seq, "+", qual, sep="\n", file=want)
        assert out[:need.value].tobytes().decode() == want.getvalue()
        bin_bytes = np.zeros(2, dtype=np.int64); bin_off = np.zeros(2, dtype=np.int64)
        rargs = (vp(arr), vp(recs), vp(results), vp(label), n, 2, 1, 1, 30, vp(bin_bytes))
        assert lib.qcb_format_records(*rargs, None, 0, vp(bin_off), vp(kept), 3) == 0
        out = np.zeros(int(bin_bytes.sum()) + 1, dtype=np.uint8)
        assert lib.qcb_format_records(*rargs, vp(out), int(out.size), vp(bin_off), vp(kept), 3) == 0
        for b in range(2):
            want = io.StringIO()
            for i, (title, seq, qual) in enumerate(reads):
                if label[i] != b:
                    continue
                seq, qual = seq[results["trim5p"][i]:results["trim3p"][i]], qual[results["trim5p"][i]:results["trim3p"][i]]
                if len(seq) < 30:
                    continue
                cols = title.replace("\t", " ").split(" ")
                name, comment = cols[0], (" ".join(cols[1:]) if len(cols) > 1 else None)
                print("@" + name + " " + (comment or ""), seq, "+", qual, sep="\n", file=want)
            assert out[bin_off[b]:bin_off[b] + bin_bytes[b]].tobytes().decode() == want.getvalue()


def test_reader_streams_from_stdin_and_fifos(tmp_path):
    """path "-" = stdin (the CLI's default input, cli.py:256-259) and any non-regular file are read sequentially: same
    chunks as from a regular file, including a record split across reads and a missing final newline."""
    import subprocess
    import sys
    reads = _big_reads(1500, 21, "I5#@+")
    data = _fastq_bytes(reads)[:-1]
    path = tmp_path / "r.fastq"
    path.write_bytes(data)
    child = ("import sys; sys.path.insert(0, %r); from qcat_b200 import fastx\n"
             "rd = fastx.Reader('-', 200000, threads=3); total = 0; lens = 0\n"
             "for ch in rd.chunks(128):\n"
             "    total += len(ch); lens += int(ch.recs['seq_len'].sum()); ch.release()\n"
             "rd.close(); print(total, lens)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # through a pipe that delivers small pieces
    feeder = subprocess.Popen([sys.executable, "-c", "import sys,time\nd=open(%r,'rb').read()\nfor i in range(0,len(d),70001):\n    sys.stdout.buffer.write(d[i:i+70001]); sys.stdout.buffer.flush()" % str(path)],
                              stdout=subprocess.PIPE)
    out = subprocess.run([sys.executable, "-c", child], stdin=feeder.stdout, capture_output=True, text=True, timeout=120)
    feeder.wait()
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == [str(len(reads)), str(sum(len(s) for _, s, _ in reads))]
    # redirected regular file on stdin takes the pread path
    with open(path, "rb") as fh:
        out = subprocess.run([sys.executable, "-c", child], stdin=fh, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.split()[0] == str(len(reads)), out.stderr


def test_four_bit_window_packers_agree_with_a_numpy_model(tmp_path):
    """qcb_pack_windows4 (records -> 4-bit windows) and qcb_pack_ascii4 (ASCII windows -> 4-bit windows): two classes per
    byte, even position in the low nibble, zero beyond the window; wrapped records included.  Host only."""
    import ctypes
    from qcat_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(21)
    cls = rng.integers(0, 16, size=256).astype(np.uint8)
    seqs = ["".join(rng.choice(list("ACGTNacgtRYX-"), size=int(n))) for n in rng.integers(0, 400, size=120)]
    text = "".join(">r%d\n%s" % (i, "".join(s[j:j + 61] + "\n" for j in range(0, len(s), 61)) or "\n") for i, s in enumerate(seqs))
    buf = text.encode("latin-1")
    recs, consumed, fastq = fastx.index_buffer(buf)
    assert len(recs) == len(seqs) and not fastq
    W = 150
    win5p, tail3p, wlen, read_len = fastx.pack_windows(buf, recs, W, threads=3, classes=cls)
    assert win5p.shape == (len(seqs), 80)
    win5, tail3, wlen_a, read_len_a = fastx.pack_windows(buf, recs, W, threads=3)
    np.testing.assert_array_equal(wlen, wlen_a)
    np.testing.assert_array_equal(read_len, read_len_a)

    def model(windows):
        out = np.zeros((len(windows), 80), dtype=np.uint8)
        for i, row in enumerate(windows):
            k = int(wlen[i])
            c = np.zeros(160, dtype=np.uint8)
            c[:k] = cls[row[:k]]
            out[i] = c[0::2] | (c[1::2] << 4)
        return out

    np.testing.assert_array_equal(win5p, model(win5))
    np.testing.assert_array_equal(tail3p, model(tail3))
    for windows, want in ((win5, win5p), (tail3, tail3p)):
        got = np.zeros((len(seqs), 80), dtype=np.uint8)
        rc = lib.qcb_pack_ascii4(ctypes.c_void_p(windows.ctypes.data), 160, ctypes.c_void_p(wlen.ctypes.data), len(seqs),
                                 ctypes.c_void_p(cls.ctypes.data), ctypes.c_void_p(got.ctypes.data), 80, 2)
        assert rc == 0
        np.testing.assert_array_equal(got, want)
