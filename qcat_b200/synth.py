"""Seeded synthetic read generator.  Only what the detection path looks at is materialised:
the first and last `W` bases of every read plus its length -- exactly the buffers qcb_detect takes.

Model per read: 10 % carry no adapter, 2 % carry a layout of a different kit; otherwise barcode b ~ U(set);
the 5' end is U{0..40} random bases + a 5'-capable layout with b filled in + insert (present in 95 % of
barcoded reads), the 3' end is insert + revcomp(layout with b) + U{0..20} random bases (70 %; 1 % of reads get a
different barcode there -> end conflict).  Adapter copies pass through an i.i.d. error channel (8 %
substitution, 6 % deletion, 5 % insertion by default); 0.1 % of all bases become N.  Read length is log-normal
(mean ~8 kb) clipped to [300, 50000].  Everything is vectorised numpy on a `numpy.random.Generator(PCG64(seed))`.
"""
import numpy as np

from qcat_b200.tables import COMPLEMENT

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _noisy_copies(rng, templates, lengths, sub, dele, ins):
    """Apply the error channel to each row of `templates` (uint8 [n, Lmax], valid up to lengths[i]).
    Returns (out uint8 [n, 2*Lmax], out_len int32 [n])."""
    n, lmax = templates.shape
    u = rng.random((n, lmax), dtype=np.float32)
    valid = np.arange(lmax)[None, :] < lengths[:, None]
    is_del = (u < dele) & valid
    is_sub = (u >= dele) & (u < dele + sub) & valid
    is_ins = (u >= dele + sub) & (u < dele + sub + ins) & valid
    emit = np.where(valid, 1, 0) - is_del.astype(np.int32) + is_ins.astype(np.int32)
    start = np.cumsum(emit, axis=1) - emit
    out_len = emit.sum(axis=1).astype(np.int32)
    out = np.zeros((n, 2 * lmax), dtype=np.uint8)
    base = templates.copy()
    # substitution: rotate within ACGT by 1..3 (N placeholders never reach here: barcodes are filled in)
    code = np.searchsorted(_ACGT, np.minimum(base, 84))          # A,C,G,T -> 0..3 (others clamp, unused)
    shifted = _ACGT[(code + rng.integers(1, 4, size=base.shape)) % 4]
    base = np.where(is_sub, shifted, base)
    rows = np.repeat(np.arange(n), lmax).reshape(n, lmax)
    keep = valid & ~is_del
    out[rows[keep], start[keep]] = base[keep]
    extra = _ACGT[rng.integers(0, 4, size=base.shape)]
    out[rows[is_ins], start[is_ins] + 1] = extra[is_ins]
    return out, out_len


def _place(windows, noisy, noisy_len, offset, active):
    """windows[i, offset[i] : offset[i] + noisy_len[i]] = noisy[i] (clipped to the window) where active."""
    n, W = windows.shape
    cols = np.arange(noisy.shape[1])[None, :]
    dst = offset[:, None] + cols
    mask = active[:, None] & (cols < noisy_len[:, None]) & (dst >= 0) & (dst < W)
    rows = np.broadcast_to(np.arange(n)[:, None], dst.shape)
    windows[rows[mask], dst[mask]] = noisy[mask]


def _filled(layout, b1, b2=None):
    seq = layout.sequence
    p1 = layout.barcode_pos_1
    out = seq[:p1.start] + b1 + seq[p1.end + 1:]
    if b2 is not None and layout.barcode_pos_2.end > -1:
        p2 = layout.barcode_pos_2
        out = out[:p2.start] + b2 + out[p2.end + 1:]
    return out


def generate(layouts, n_reads, seed, W=150, stride=160, foreign_layouts=(), sub=0.08, dele=0.06, ins=0.05,
             p_none=0.10, p_foreign=0.02, p5=0.95, p3=0.70, p_conflict=0.01, p_n=0.001, mean_len=8000.0,
             min_len=300, max_len=50000, length_model="lognormal"):
    """Returns dict(win5, tail3, wlen, read_len, truth_barcode, truth_layout5, truth_layout3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(n_reads)
    layouts = list(layouts)
    nl = len(layouts)
    dual = all(l.barcode_set_2 is not None for l in layouts)
    # template table: tmpl[layout][barcode] -> bytes
    n1 = [len(l.barcode_set_1) for l in layouts]
    lmax = max(len(l.sequence) for l in layouts + list(foreign_layouts))
    table = []
    for l in layouts:
        rows = np.zeros((len(l.barcode_set_1), lmax), dtype=np.uint8)
        for i, b in enumerate(l.barcode_set_1):
            s = _filled(l, b.sequence).encode()
            rows[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        table.append(rows)
    foreign_rows = []
    for l in foreign_layouts:
        b = l.barcode_set_1[0].sequence
        b2 = l.barcode_set_2[0].sequence if l.barcode_set_2 else None
        s = _filled(l, b, b2).encode()
        row = np.zeros(lmax, dtype=np.uint8)
        row[:len(s)] = np.frombuffer(s, dtype=np.uint8)
        foreign_rows.append((row, len(s)))

    kind = rng.random(n)
    is_none = kind < p_none
    is_foreign = (kind >= p_none) & (kind < p_none + p_foreign) & (len(foreign_rows) > 0)
    barcoded = ~is_none & ~is_foreign

    lay5 = rng.integers(0, nl, size=n)
    lay3 = rng.integers(0, nl, size=n)
    nb = np.array(n1)
    bc5 = (rng.random(n) * nb[lay5]).astype(np.int64)
    bc3 = np.minimum(bc5, nb[lay3] - 1)
    conflict = barcoded & (rng.random(n) < p_conflict)
    bc3 = np.where(conflict, (bc3 + 1 + (rng.random(n) * (nb[lay3] - 1)).astype(np.int64)) % nb[lay3], bc3)
    has5 = (barcoded & (rng.random(n) < p5)) | is_foreign
    has3 = barcoded & (rng.random(n) < p3)
    has3 |= barcoded & ~has5                                           # at least one end

    def templates_for(lay, bc, foreign_mask):
        t = np.zeros((n, lmax), dtype=np.uint8)
        tl = np.zeros(n, dtype=np.int32)
        for li in range(nl):
            m = lay == li
            if m.any():
                t[m] = table[li][bc[m]]
                tl[m] = len(layouts[li].sequence)
        if foreign_rows and foreign_mask.any():
            pick = rng.integers(0, len(foreign_rows), size=n)
            for fi, (row, flen) in enumerate(foreign_rows):
                m = foreign_mask & (pick == fi)
                t[m] = row
                tl[m] = flen
        if dual:
            # second barcode of dual layouts: independent uniform draw per read, written over the 2nd placeholder
            for li, l in enumerate(layouts):
                m = np.nonzero(lay == li)[0]
                if m.size == 0:
                    continue
                p2 = l.barcode_pos_2
                set2 = np.stack([np.frombuffer(b.sequence.encode(), dtype=np.uint8) for b in l.barcode_set_2])
                pick2 = rng.integers(0, len(l.barcode_set_2), size=m.size)
                t[m[:, None], np.arange(p2.start, p2.end + 1)[None, :]] = set2[pick2]
        return t, tl

    win5 = _ACGT[rng.integers(0, 4, size=(n, W))]
    tail3 = _ACGT[rng.integers(0, 4, size=(n, W))]

    t5, tl5 = templates_for(lay5, bc5, is_foreign)
    noisy5, nlen5 = _noisy_copies(rng, t5, tl5, sub, dele, ins)
    _place(win5, noisy5, nlen5, rng.integers(0, 41, size=n).astype(np.int64), has5)

    t3, tl3 = templates_for(lay3, bc3, np.zeros(n, dtype=bool))
    noisy3, nlen3 = _noisy_copies(rng, t3, tl3, sub, dele, ins)
    # reverse complement each noisy copy in place (left-aligned), then right-align it r bases before the read end
    cols = np.arange(noisy3.shape[1])[None, :]
    src = np.clip(nlen3[:, None] - 1 - cols, 0, noisy3.shape[1] - 1)
    rc3 = COMPLEMENT[np.take_along_axis(noisy3, src, axis=1)]
    rc3 = np.where(cols < nlen3[:, None], rc3, 0).astype(np.uint8)
    r = rng.integers(0, 21, size=n).astype(np.int64)
    _place(tail3, rc3, nlen3, W - r - nlen3.astype(np.int64), has3)

    for w in (win5, tail3):
        w[rng.random(w.shape, dtype=np.float32) < p_n] = ord("N")

    if length_model == "loguniform":          # "mixed lengths": every octave of [min_len, max_len] equally likely
        lo = float(max(min_len, 2 * W))
        read_len = np.floor(lo * np.exp(rng.random(n) * np.log(float(max_len) / lo))).astype(np.int64)
        read_len = np.clip(read_len, int(lo), max_len)
    else:
        sigma = 0.9
        mu = np.log(mean_len) - 0.5 * sigma * sigma
        read_len = np.clip(rng.lognormal(mu, sigma, size=n), max(min_len, 2 * W), max_len).astype(np.int64)

    out5 = np.zeros((n, stride), dtype=np.uint8)
    out3 = np.zeros((n, stride), dtype=np.uint8)
    out5[:, :W] = win5
    out3[:, :W] = tail3
    truth = np.where(barcoded, bc5, -1).astype(np.int32)
    return {"win5": out5, "tail3": out3, "wlen": np.full(n, W, dtype=np.int32), "read_len": read_len,
            "truth_barcode": truth, "truth_conflict": conflict, "has5": has5, "has3": has3,
            "layout5": lay5.astype(np.int32), "layout3": lay3.astype(np.int32)}


def windows_to_reads(data, indices=None):
    """Materialise full read strings (window5 + filler + tail3) for the reference's string API; filler is 'A's so
    the read has the stated length.  Only valid for reads with read_len >= 2 * W."""
    W = int(data["wlen"][0])
    idx = range(len(data["wlen"])) if indices is None else indices
    reads = []
    for i in idx:
        n = int(data["read_len"][i])
        head = bytes(data["win5"][i, :W]).decode("latin-1")
        tail = bytes(data["tail3"][i, :W]).decode("latin-1")
        reads.append(head + "A" * (n - 2 * W) + tail if n >= 2 * W else (head + tail)[:max(n, W)])
    return reads


_PARALLEL_JOB = None


def _parallel_chunk(index):
    layouts, chunk, seed, kwargs, n_total = _PARALLEL_JOB
    lo = index * chunk
    base = [int(v) for v in np.atleast_1d(seed)]
    return generate(layouts, min(chunk, n_total - lo), seed=base + [int(index)], **kwargs)


def generate_parallel(layouts, n_reads, seed, workers=None, chunk=65536, **kwargs):
    """`generate` for large batches: the reads are drawn in chunks of `chunk` reads, chunk c from the generator seeded
    with (seed, c), by a pool of forked worker processes -- the result depends on (seed, chunk) only, not on the worker
    count.  Fork-based: call it before the process initialises CUDA (workers only run numpy)."""
    import multiprocessing
    import os
    global _PARALLEL_JOB
    n = int(n_reads)
    n_chunks = max(1, (n + chunk - 1) // chunk)
    workers = max(1, min(n_chunks, int(workers or os.cpu_count() or 1)))
    _PARALLEL_JOB = (list(layouts), int(chunk), seed, kwargs, n)
    try:
        if workers == 1:
            parts = [_parallel_chunk(i) for i in range(n_chunks)]
        else:
            with multiprocessing.get_context("fork").Pool(workers) as pool:
                parts = pool.map(_parallel_chunk, range(n_chunks))
    finally:
        _PARALLEL_JOB = None
    return {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=0)) for k in parts[0]}
