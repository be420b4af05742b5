"""Native FASTQ / FASTA ingest and per-barcode output around the hot path.

Mirrors what `qcat/cli.py` does around `detect_barcode_batch` -- `iter_fastx` (cli.py:235-306), trimming and the
min-length filter (:521-530), the TSV lines (:408-442) and the per-barcode files of `-b` (:309-358) -- but on memory
buffers: records are indexed and cut into windows by libqcat_b200.so (`qcb_fastx_index`, `qcb_pack_windows`), scored
in batches of `batch_size` reads exactly like the CLI (the kit vote is per batch), and written through
`qcb_format_records`.  Strings are only materialised for the TSV.
"""
import ctypes
import os

import numpy as np

from qcat_b200 import _ffi


class FastxError(ValueError):
    """Malformed input, like the ValueError the reference's Bio parsers raise."""


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def _check_io(rc):
    if rc != 0:
        msg = _ffi.load().qcb_io_last_error()
        raise FastxError(msg.decode("utf-8", "replace") if msg else "fastx error")


def index_buffer(buf, final_chunk=True, max_records=None):
    """Index the complete records of a bytes-like FASTQ / FASTA chunk.
    Returns (records structured array, bytes consumed, is_fastq)."""
    lib = _ffi.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    if max_records is None:
        max_records = max(16, int(arr.size // 12) + 16)          # a record needs more than 12 bytes
    recs = np.zeros(max_records, dtype=_ffi.RECORD_DTYPE)
    n = ctypes.c_int64(0)
    consumed = ctypes.c_int64(0)
    fastq = ctypes.c_int32(1)
    _check_io(lib.qcb_fastx_index(_vp(arr) if arr.size else None, int(arr.size), 1 if final_chunk else 0, _vp(recs),
                                  int(max_records), ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(fastq))
              if arr.size else 0)
    return recs[:n.value], int(consumed.value), bool(fastq.value)


def pack_windows(buf, recs, max_align_length=150, threads=None):
    """(win5, tail3, wlen, read_len) of indexed records -- the buffers DevicePlan.detect takes."""
    lib = _ffi.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    n = len(recs)
    W = int(max_align_length)
    stride = max(16, (W + 15) // 16 * 16)
    win5 = np.empty((n, stride), dtype=np.uint8)
    tail3 = np.empty((n, stride), dtype=np.uint8)
    wlen = np.empty(n, dtype=np.int32)
    read_len = np.empty(n, dtype=np.int64)
    recs = np.ascontiguousarray(recs)
    if n:
        _check_io(lib.qcb_pack_windows(_vp(arr), _vp(recs), n, W, stride, _vp(win5), _vp(tail3), _vp(wlen), _vp(read_len),
                                       threads or os.cpu_count() or 1))
    return win5, tail3, wlen, read_len


def iter_chunks(path, chunk_bytes=64 << 20, multiple_of=1):
    """Yield (buffer, records, is_fastq) for consecutive chunks of a FASTQ / FASTA file; records never straddle chunks
    and every chunk but the last holds a multiple of `multiple_of` records (so CLI batches of 4000 stay aligned)."""
    carry = b""
    with open(path, "rb") as fh:
        while True:
            block = fh.read(chunk_bytes)
            final = len(block) < chunk_bytes
            buf = carry + block
            if not buf:
                return
            recs, consumed, fastq = index_buffer(buf, final_chunk=final)
            if not final and multiple_of > 1:
                keep = (len(recs) // multiple_of) * multiple_of
                if keep < len(recs):
                    consumed = int(recs[keep]["title_off"]) - 1          # start of the first record carried over
                    recs = recs[:keep]
            if len(recs):
                yield buf, recs, fastq
            carry = buf[consumed:]
            if final:
                if carry.strip():
                    raise FastxError("trailing bytes that do not form a record")
                return


def _title(buf, rec):
    header = bytes(buf[int(rec["title_off"]):int(rec["title_off"] + rec["title_len"])]).decode("latin-1")
    cols = header.replace("\t", " ").split(" ")                  # extract_fastx_comment, cli.py:199-213
    return cols[0], (" ".join(cols[1:]) if len(cols) > 1 else None)


def demux_file(path, scanner, qcat_config=None, batch_size=4000, trim=False, min_read_length=0, out_dir=None, tsv=None,
               nobatch=False, chunk_bytes=64 << 20):
    """Demultiplex a FASTQ / FASTA file like `qcat -f path [-b out_dir] [--tsv] [--trim]` (cli.py:445-563).

    scanner: a qcat_b200 (or drop-in patched qcat) scanner.  Reads are scored in batches of `batch_size` (4000 in the
    CLI; `nobatch` = single-read mode without the kit vote).  out_dir: per-barcode files as with `-b`; tsv: a text
    file object receiving the `--tsv` table.  Returns {"reads", "skipped", "barcodes": {name: count}, "records"}.
    """
    from qcat_b200 import scanner as qscanner
    qcat_config = qcat_config or qscanner._default_config()
    lib = _ffi.load()
    plan = scanner._plan_for(qcat_config)
    tables = plan.tables
    threads = os.cpu_count() or 1
    files = {}
    counts = {}
    all_records = []
    total = skipped = 0
    if tsv is not None:
        print("name", "length", "barcode", "score", "kit", "adapter_end", "comment", sep="\t", file=tsv)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)

    # output bins: 0 = none, then one per distinct barcode name seen (dual: names are synthesised per pair)
    bin_names = ["none"]
    bin_of_name = {"none": 0}

    for buf, recs, fastq in iter_chunks(path, chunk_bytes, multiple_of=1 if nobatch else batch_size):
        arr = np.frombuffer(buf, dtype=np.uint8)
        win5, tail3, wlen, read_len = pack_windows(buf, recs, qcat_config.max_align_length, threads)
        n = len(recs)
        results = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        names = [layout.kit for layout in scanner.layouts]
        step = 1 if nobatch else batch_size
        if nobatch or len(set(names)) == 1:
            # no vote needed (single read mode, or every layout names the same kit): one device call per chunk
            kits = scanner.layouts
            subset = scanner._subset_for(plan, kits)
            plan.detect(win5, tail3, wlen, read_len, subset, out=results)
        else:
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                vote = plan.kit_vote(win5[lo:hi], tail3[lo:hi], wlen[lo:hi])
                kit = scanner._kit_from_votes(vote, names)
                subset = tables.kit_subset(kit)
                plan.detect(win5[lo:hi], tail3[lo:hi], wlen[lo:hi], read_len[lo:hi], subset, out=results[lo:hi])
        if getattr(scanner, "enable_filter_barcodes", False) and not nobatch:
            for lo in range(0, n, step):                           # filter_barcodes works per CLI batch
                hi = min(n, lo + step)
                dicts = [scanner._record_to_dict(plan, r) for r in results[lo:hi]]
                count = {}
                for d in dicts:
                    scanner.update_barcode_count(d, count)
                valid = scanner.get_valid(count, 0.05)
                for i, d in enumerate(dicts):
                    if d["barcode"] and d["barcode"].id not in valid:
                        results[lo + i] = (-1, -1, 0.0, 0, results[lo + i]["trim5p"], results[lo + i]["trim3p"], 1)
        all_records.append(results)

        # barcode name per record (vectorised through a small lookup over the distinct (layout, barcode) pairs)
        key = results["layout"].astype(np.int64) * (1 << 32) + results["barcode"].astype(np.int64)
        called = (results["layout"] >= 0) & (results["barcode"] >= 0)
        bins = np.zeros(n, dtype=np.int32)
        for k in np.unique(key[called]):
            layout_index, barcode_index = int(k >> 32), int(k & 0xffffffff)
            b = tables.barcode_object(layout_index, barcode_index)
            name = "barcode{:02d}/{:02d}".format(b[0].id, b[1].id) if tables.mode == 1 else b.name
            if name not in bin_of_name:
                bin_of_name[name] = len(bin_names)
                bin_names.append(name)
            bins[called & (key == k)] = bin_of_name[name]

        # lengths after trimming and the min-length filter (cli.py:521-530)
        if trim:
            a = np.clip(results["trim5p"].astype(np.int64), 0, read_len)
            b = np.maximum(np.clip(results["trim3p"].astype(np.int64), 0, read_len), a)
            out_len = b - a
        else:
            out_len = read_len
        kept = out_len >= min_read_length
        total += n
        skipped += int((~kept).sum())
        for b_index, cnt in zip(*np.unique(bins[kept], return_counts=True)):
            counts[bin_names[b_index]] = counts.get(bin_names[b_index], 0) + int(cnt)

        if tsv is not None:
            for i in np.nonzero(kept)[0]:
                name, comment = _title(arr, recs[i])
                r = results[i]
                if bins[i]:
                    b = tables.barcode_object(int(r["layout"]), int(r["barcode"]))
                    bid = "{}/{}".format(b[0].id, b[1].id) if tables.mode == 1 else b.id
                    print(name, int(out_len[i]), bid, float(r["barcode_score"]), tables.layouts[int(r["layout"])].kit,
                          int(r["adapter_end"]), comment, sep="\t", file=tsv)
                else:
                    print(name, int(out_len[i]), "none", "-1", "none", "-1", comment, sep="\t", file=tsv)

        if out_dir:
            n_bins = len(bin_names)
            bin_bytes = np.zeros(n_bins, dtype=np.int64)
            bin_off = np.zeros(n_bins, dtype=np.int64)
            kept_u8 = np.zeros(n, dtype=np.uint8)
            recs_c = np.ascontiguousarray(recs)
            args = (_vp(arr), _vp(recs_c), _vp(results), _vp(bins), n, n_bins, 1 if fastq else 0, 1 if trim else 0,
                    int(min_read_length), _vp(bin_bytes))
            _check_io(lib.qcb_format_records(*args, None, 0, _vp(bin_off), _vp(kept_u8), threads))
            out = np.empty(int(bin_bytes.sum()) + 1, dtype=np.uint8)
            _check_io(lib.qcb_format_records(*args, _vp(out), int(out.size), _vp(bin_off), _vp(kept_u8), threads))
            for b_index in range(n_bins):
                if bin_bytes[b_index] == 0:
                    continue
                name = bin_names[b_index].replace("/", "_")
                if name not in files:
                    files[name] = open(os.path.join(out_dir, name + (".fastq" if fastq else ".fasta")), "wb")
                files[name].write(out[bin_off[b_index]:bin_off[b_index] + bin_bytes[b_index]].tobytes())
    for fh in files.values():
        fh.close()
    records = np.concatenate(all_records) if all_records else np.zeros(0, dtype=_ffi.RESULT_DTYPE)
    return {"reads": total, "skipped": skipped, "barcodes": counts, "records": records}
