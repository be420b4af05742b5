"""CPU: property tests of the native record scanner on small adversarial inputs.

* the parallel scan (forced with threads < 0) returns byte-identical records -- or the same error -- as the serial scan;
* the serial scan agrees with a pure-Python model of Bio.SeqIO.QualityIO.FastqGeneralIterator /
  Bio.SeqIO.FastaIO.SimpleFastaParser (what qcat/cli.py:235-306 iterates with) on well-formed input.
"""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from qcat_b200 import fastx

SEQ = st.text(alphabet="ACGTNacgtn", min_size=0, max_size=90)
QUAL_CHARS = "@+!#5I~>\""                     # '@' and '+' first: the classic ambiguity of the format
TITLE = st.text(alphabet="abcXYZ019_=:/ \t@+>", min_size=1, max_size=24).map(lambda t: t.strip() or "r").filter(
    lambda t: t[0] not in " \t")


@st.composite
def fastq_records(draw):
    n = draw(st.integers(min_value=0, max_value=40))
    recs = []
    for _ in range(n):
        seq = draw(SEQ)
        qual = "".join(draw(st.lists(st.sampled_from(QUAL_CHARS), min_size=len(seq), max_size=len(seq))))
        recs.append((draw(TITLE), seq, qual))
    return recs


def render_fastq(recs, newline, wrap, blank_lines, trailing_newline):
    out = []
    for title, seq, qual in recs:
        if wrap and len(seq) > wrap:
            s_lines = [seq[i:i + wrap] for i in range(0, len(seq), wrap)]
            q_lines = [qual[i:i + wrap] for i in range(0, len(qual), wrap)]
        else:
            s_lines, q_lines = [seq], [qual]
        out.append("@" + title + newline + newline.join(s_lines) + newline + "+" + newline + newline.join(q_lines) + newline)
        if blank_lines:
            out.append(newline)
    text = "".join(out)
    if not trailing_newline and text.endswith(newline):
        text = text[:-len(newline)]
    return text.encode("latin-1")


def check_against_model(buf, recs, got):
    assert len(got) == len(recs)
    for r, (title, seq, qual) in zip(got, recs):
        assert buf[r["title_off"]:r["title_off"] + r["title_len"]].decode("latin-1") == title.rstrip()
        assert int(r["seq_len"]) == len(seq)
        seq_text = buf[r["seq_off"]:r["seq_off"] + r["seq_span"]].decode("latin-1")
        assert "".join(l.rstrip() for l in seq_text.split("\n")) == seq
        if r["qual_off"] >= 0:
            qual_text = buf[r["qual_off"]:r["qual_off"] + r["qual_span"]].decode("latin-1")
            assert "".join(l.rstrip() for l in qual_text.split("\n")) == qual


def same_outcome(buf, final_chunk, threads):
    outcomes = []
    for t in (1, threads):
        try:
            recs, consumed, is_fastq = fastx.index_buffer(buf, final_chunk=final_chunk, threads=t)
            outcomes.append(("ok", recs.tobytes(), consumed, is_fastq))
        except fastx.FastxError as exc:
            outcomes.append(("error",))
    assert outcomes[0] == outcomes[1], "serial and parallel scans disagree"
    return outcomes[0]


@settings(deadline=None, max_examples=300, suppress_health_check=[HealthCheck.too_slow])
@given(recs=fastq_records(), crlf=st.booleans(), blank_lines=st.booleans(), trailing_newline=st.booleans(),
       threads=st.integers(min_value=2, max_value=9))
def test_four_line_fastq(recs, crlf, blank_lines, trailing_newline, threads):
    recs = [(t, s, q) for t, s, q in recs if not (blank_lines and not s)]      # an empty record next to blank lines is ambiguous
    buf = render_fastq(recs, "\r\n" if crlf else "\n", 0, blank_lines, trailing_newline)
    outcome = same_outcome(buf, True, -threads)
    assert outcome[0] == "ok"
    got = np.frombuffer(outcome[1], dtype=fastx._ffi.RECORD_DTYPE)
    check_against_model(buf, recs, got)
    assert outcome[2] == len(buf) or not buf.strip()
    pack = fastx.pack_windows(buf, got, 50, threads=2)
    for i, (_, seq, _) in enumerate(recs):
        k = min(len(seq), 50)
        assert pack[2][i] == k and pack[3][i] == len(seq)
        assert pack[0][i, :k].tobytes().decode() == seq[:k] and pack[1][i, :k].tobytes().decode() == seq[len(seq) - k:]


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(recs=fastq_records(), wrap=st.integers(min_value=3, max_value=40), threads=st.integers(min_value=2, max_value=9),
       cut=st.integers(min_value=0, max_value=400))
def test_wrapped_fastq_and_partial_chunks(recs, wrap, threads, cut):
    """Wrapped records defeat the local record-start test, so the parallel scan must fall back; a chunk that ends in the
    middle of a record (final_chunk=False) must stop at the same record in both scans."""
    recs = [(t, s, q) for t, s, q in recs if s and not (q[:1] in "@+" or any(q[i] in "@+" for i in range(0, len(q), wrap)))]
    buf = render_fastq(recs, "\n", wrap, False, True)
    outcome = same_outcome(buf, True, -threads)
    assert outcome[0] == "ok"
    check_against_model(buf, recs, np.frombuffer(outcome[1], dtype=fastx._ffi.RECORD_DTYPE))
    part = buf[:max(0, len(buf) - cut)]
    same_outcome(part, False, -threads)


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(data=st.binary(min_size=0, max_size=300), threads=st.integers(min_value=2, max_value=9), final=st.booleans())
def test_garbage_never_disagrees(data, threads, final):
    """Arbitrary bytes, with and without a leading '@' / '>': both scans end the same way (records or error), no crash."""
    for prefix in (b"", b"@", b">", b"@a\nAC\n+\nII\n"):
        same_outcome(prefix + data, final, -threads)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(recs=st.lists(st.tuples(TITLE.filter(lambda t: ">" not in t), SEQ), min_size=0, max_size=30),
       wrap=st.integers(min_value=5, max_value=60), threads=st.integers(min_value=2, max_value=9))
def test_fasta(recs, wrap, threads):
    text = "".join(">%s\n%s" % (t, "".join(s[i:i + wrap] + "\n" for i in range(0, len(s), wrap)) or "\n") for t, s in recs)
    buf = text.encode("latin-1")
    outcome = same_outcome(buf, True, -threads)
    assert outcome[0] == "ok"
    got = np.frombuffer(outcome[1], dtype=fastx._ffi.RECORD_DTYPE)
    check_against_model(buf, [(t, s, "") for t, s in recs], got)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(recs=st.lists(st.tuples(TITLE.filter(lambda t: ">" not in t),
                               st.text(alphabet="ACGTN \r", min_size=0, max_size=80)), min_size=1, max_size=20),
       wrap=st.integers(min_value=5, max_value=60), threads=st.integers(min_value=2, max_value=9))
def test_fasta_interior_blanks(recs, wrap, threads):
    """Blanks and carriage returns INSIDE FASTA sequence lines are not bases: Bio's SimpleFastaParser (what cli.py:235-306
    iterates with) joins line.rstrip() pieces and then removes every ' ' and '\\r'."""
    text = "".join(">%s\n%s" % (t, "".join(s[i:i + wrap] + "\n" for i in range(0, len(s), wrap)) or "\n") for t, s in recs)
    buf = text.encode("latin-1")
    outcome = same_outcome(buf, True, -threads)
    assert outcome[0] == "ok"
    got = np.frombuffer(outcome[1], dtype=fastx._ffi.RECORD_DTYPE)
    assert len(got) == len(recs)
    want = []
    for _, s in recs:
        lines = [s[i:i + wrap] for i in range(0, len(s), wrap)]
        want.append("".join(l.rstrip() for l in lines).replace(" ", "").replace("\r", ""))
    pack = fastx.pack_windows(buf, got, 30, threads=2)
    for i, seq in enumerate(want):
        k = min(len(seq), 30)
        assert int(got[i]["seq_len"]) == len(seq)
        assert pack[2][i] == k and pack[3][i] == len(seq)
        assert pack[0][i, :k].tobytes().decode() == seq[:k] and pack[1][i, :k].tobytes().decode() == seq[len(seq) - k:]
