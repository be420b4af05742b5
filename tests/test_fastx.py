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
