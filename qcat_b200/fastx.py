"""Native FASTQ / FASTA ingest and output around the hot path.

Mirrors what `qcat/cli.py` does around `detect_barcode_batch` -- `iter_fastx` (cli.py:235-306), trimming and the
min-length filter (:521-530), the TSV lines (:408-442), the per-barcode files of `-b` (:309-336) and the single
output stream with `barcode=<id>` comments (:337-352) -- but on memory buffers: chunks of the file are read and
indexed by several threads inside libqcat_b200.so (`qcb_reader_*`, `qcb_fastx_index_mt`), cut into windows
(`qcb_pack_windows`), scored in batches of `batch_size` reads exactly like the CLI (the kit vote is per batch), and
written through `qcb_format_records` / `qcb_format_stream` / `qcb_format_tsv`.  `demux_file` runs the three stages
(read + index + pack, device scoring, format + write) as a pipeline on three threads; the native calls release
the GIL.  No Python string is created per read.
"""
import ctypes
import io
import os
import queue
import sys
import threading

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


def _threads(threads):
    return int(threads or os.cpu_count() or 1)


def index_buffer(buf, final_chunk=True, max_records=None, threads=1):
    """Index the complete records of a bytes-like FASTQ / FASTA chunk (threads > 1: parallel scan, same result).
    Returns (records structured array, bytes consumed, is_fastq)."""
    lib = _ffi.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    n = ctypes.c_int64(0)
    consumed = ctypes.c_int64(0)
    fastq = ctypes.c_int32(1)
    if not arr.size:
        return np.zeros(0, dtype=_ffi.RECORD_DTYPE), 0, True
    # without a limit from the caller the record array grows until the scan is no longer bounded by it
    capacity = int(max_records) if max_records is not None else max(1024, int(arr.size // 64))
    while True:
        recs = np.zeros(capacity, dtype=_ffi.RECORD_DTYPE)
        _check_io(lib.qcb_fastx_index_mt(_vp(arr), int(arr.size), 1 if final_chunk else 0, _vp(recs), capacity,
                                         ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(fastq), int(threads)))
        if max_records is not None or n.value < capacity:
            return recs[:n.value], int(consumed.value), bool(fastq.value)
        capacity *= 4


def pack_windows(buf, recs, max_align_length=150, threads=None, classes=None):
    """(win5, tail3, wlen, read_len) of indexed records -- the buffers DevicePlan.detect takes; with `classes` (a plan's
    base_classes() table) the windows come in the 4-bit form DevicePlan.detect4 takes (two bases per byte)."""
    lib = _ffi.load()
    arr = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
    n = len(recs)
    W = int(max_align_length)
    if classes is not None:
        stride4 = max(8, ((W + 1) // 2 + 7) // 8 * 8)
        win5p = np.empty((n, stride4), dtype=np.uint8)
        tail3p = np.empty((n, stride4), dtype=np.uint8)
        wlen = np.empty(n, dtype=np.int32)
        read_len = np.empty(n, dtype=np.int64)
        recs = np.ascontiguousarray(recs)
        cls = np.ascontiguousarray(classes, dtype=np.uint8)
        if n:
            _check_io(lib.qcb_pack_windows4(_vp(arr), _vp(recs), n, W, stride4, _vp(cls), _vp(win5p), _vp(tail3p), _vp(wlen),
                                            _vp(read_len), _threads(threads)))
        return win5p, tail3p, wlen, read_len
    stride = max(16, (W + 15) // 16 * 16)
    win5 = np.empty((n, stride), dtype=np.uint8)
    tail3 = np.empty((n, stride), dtype=np.uint8)
    wlen = np.empty(n, dtype=np.int32)
    read_len = np.empty(n, dtype=np.int64)
    recs = np.ascontiguousarray(recs)
    if n:
        _check_io(lib.qcb_pack_windows(_vp(arr), _vp(recs), n, W, stride, _vp(win5), _vp(tail3), _vp(wlen), _vp(read_len),
                                       _threads(threads)))
    return win5, tail3, wlen, read_len


def iter_chunks(path, chunk_bytes=64 << 20, multiple_of=1, threads=1):
    """Yield (buffer, records, is_fastq) for consecutive chunks of a FASTQ / FASTA file; records never straddle chunks
    and every chunk but the last holds a multiple of `multiple_of` records (so CLI batches of 4000 stay aligned).
    Pure-Python chunking over `index_buffer`; `Reader` is the native equivalent demux_file uses."""
    carry = b""
    with open(path, "rb") as fh:
        while True:
            block = fh.read(chunk_bytes)
            final = len(block) < chunk_bytes
            buf = carry + block
            if not buf:
                return
            recs, consumed, fastq = index_buffer(buf, final_chunk=final, threads=threads)
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


class Chunk(object):
    """One chunk of complete records owned by the native reader: `.data` (uint8 view of the bytes), `.recs`
    (structured view of the record index), `.fastq`.  Views die with release()."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle
        length = ctypes.c_int64(0)
        n = ctypes.c_int64(0)
        fastq = ctypes.c_int32(1)
        data = lib.qcb_chunk_data(handle, ctypes.byref(length))
        recs = lib.qcb_chunk_records(handle, ctypes.byref(n), ctypes.byref(fastq))
        self.data = np.ctypeslib.as_array(ctypes.cast(data, ctypes.POINTER(ctypes.c_uint8)), shape=(length.value,))
        raw = np.ctypeslib.as_array(ctypes.cast(recs, ctypes.POINTER(ctypes.c_uint8)),
                                    shape=(n.value * _ffi.RECORD_DTYPE.itemsize,)) if n.value else np.zeros(0, np.uint8)
        self.recs = raw.view(_ffi.RECORD_DTYPE)
        self.fastq = bool(fastq.value)

    def __len__(self):
        return len(self.recs)

    def release(self):
        if self._handle:
            self._lib.qcb_chunk_release(self._handle)
            self._handle = None
            self.data = self.recs = None


class Reader(object):
    """Native chunked FASTQ / FASTA reader (qcb_reader_*): parallel pread + parallel record scan per chunk.
    path "-" (standard input), pipes and FIFOs are read sequentially."""

    def __init__(self, path, chunk_bytes=64 << 20, threads=None):
        self._lib = _ffi.load()
        self._handle = self._lib.qcb_reader_open(os.fsencode(path), int(chunk_bytes), _threads(threads))
        if not self._handle:
            msg = self._lib.qcb_io_last_error()
            raise IOError(msg.decode("utf-8", "replace") if msg else "cannot open %s" % path)

    def next_chunk(self, multiple_of=1):
        """The next Chunk, or None at the end of the file."""
        handle = ctypes.c_void_p(None)
        _check_io(self._lib.qcb_reader_next(self._handle, int(multiple_of), ctypes.byref(handle)))
        return Chunk(self._lib, handle.value) if handle.value else None

    def chunks(self, multiple_of=1):
        while True:
            chunk = self.next_chunk(multiple_of)
            if chunk is None:
                return
            yield chunk

    def close(self):
        if self._handle:
            self._lib.qcb_reader_close(self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _Labels(object):
    """Strings the writers need per record, interned: output bin (barcode name), barcode id, kit."""

    def __init__(self, tables):
        self.tables = tables
        self.strings = []
        self.index = {}
        self.none = self.intern("none")
        self.by_key = {}                      # (layout << 32 | barcode) -> (name label, id label, kit label)

    def intern(self, text):
        if text not in self.index:
            self.index[text] = len(self.strings)
            self.strings.append(text)
        return self.index[text]

    def lookup(self, key):
        if key not in self.by_key:
            layout_index, barcode_index = int(key >> 32), int(key & 0xffffffff)
            b = self.tables.barcode_object(layout_index, barcode_index)
            if self.tables.mode == 1:           # dual: names and ids are synthesised per pair (scanner_dual.py:132-136)
                name, bid = "barcode{:02d}/{:02d}".format(b[0].id, b[1].id), "{}/{}".format(b[0].id, b[1].id)
            else:
                name, bid = b.name, str(b.id)
            # simple mode: the adapter is None and the reference prints just that (cli.py:423-426: kit_name = None)
            kit = "None" if self.tables.mode == 2 else str(self.tables.layouts[layout_index].kit)
            self.by_key[key] = (self.intern(name), self.intern(bid), self.intern(kit))
        return self.by_key[key]

    def table(self):
        blobs = [s.encode("latin-1") for s in self.strings]
        off = np.zeros(len(blobs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(b) for b in blobs])
        return np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8), off


def _batch_kits(vote, kit_of_layout, n_kits, batch_size):
    """Per CLI batch: the kit named by most reads' best-scoring end; ties go to the kit seen first in the batch
    (detect_kit / get_most_abundant_kits, scanner_base.py:657-678: dict insertion order + stable sort)."""
    kits = []
    v = kit_of_layout[vote]
    for lo in range(0, len(v), batch_size):
        part = v[lo:lo + batch_size]
        counts = np.bincount(part, minlength=n_kits)
        best = np.nonzero(counts == counts.max())[0]
        if len(best) > 1:
            first = [int(np.argmax(part == k)) for k in best]
            best = [best[int(np.argmin(first))]]
        kits.append(int(best[0]))
    return kits


def _read_bytes(chunk, rec):
    """The record's sequence as Bio's parsers return it (line pieces joined; FASTA: interior blanks removed)."""
    off, span = int(rec["seq_off"]), int(rec["seq_span"])
    raw = chunk.data[off:off + span].tobytes()
    if span != int(rec["seq_len"]):
        raw = b"".join(line.rstrip() for line in raw.split(b"\n"))
        if int(rec["qual_off"]) < 0:
            raw = raw.replace(b" ", b"").replace(b"\r", b"")
    return raw


def _middle_scan_chunk(scanner, chunk, results, qcat_config):
    """--detect-middle on the records of one chunk (scanner_base.py:593-595, after the two-end decision and before the
    per-batch barcode filter): reads whose body read[W:-W] still holds an adapter of the detected kit become empty
    results with exit_status 997; trims are kept."""
    W = int(qcat_config.max_align_length)
    tables_layouts = scanner._plan_for(qcat_config).tables.layouts
    by_kit = {}
    for i in np.nonzero(results["layout"] >= 0)[0].tolist():
        by_kit.setdefault(tables_layouts[int(results["layout"][i])].kit, []).append(i)
    for kit_name, indices in by_kit.items():
        bodies = [_read_bytes(chunk, chunk.recs[i])[W:-W] for i in indices]
        found = scanner._middle_found(kit_name, bodies, qcat_config)
        for i, hit in zip(indices, found):
            if hit:
                results[i] = (-1, -1, 0.0, 0, results["trim5p"][i], results["trim3p"][i], 997)


def _score_chunk(scanner, plan, packed, batch_size, nobatch, chunk=None, qcat_config=None, four_bit=False):
    """qcb_result records of one chunk, batch semantics of the CLI loop (cli.py:500-513).  four_bit: the windows are in
    the plan's 4-bit form (pack_windows(..., classes=plan.base_classes()))."""
    win5, tail3, wlen, read_len = packed
    tables = plan.tables
    n = len(wlen)
    results = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    names = [layout.kit for layout in scanner.layouts]
    kit_names = list(dict.fromkeys(names))
    detect = plan.detect4 if four_bit else plan.detect
    if tables.mode == 2:
        detect(win5, tail3, wlen, read_len, out=results)        # simple mode: no layouts, no vote (scanner_simple.py)
    elif nobatch or len(kit_names) == 1:
        # no vote needed (single-read mode, or every layout names the same kit): one device call per chunk
        detect(win5, tail3, wlen, read_len, scanner._subset_for(plan, scanner.layouts), out=results)
    elif tables.kit_index()[1] is not None and hasattr(plan, "detect_auto"):
        # one pass (qcb_detect_auto): adapter stage over all layouts once, per-batch vote and kit restriction on the device
        (plan.detect_auto4 if four_bit else plan.detect_auto)(win5, tail3, wlen, read_len, tables.kit_index()[1], batch_size,
                                                              out=results)
    else:
        vote = plan.kit_vote(win5, tail3, wlen)                 # one device call for the whole chunk
        kit_of_layout = np.array([kit_names.index(k) for k in names], dtype=np.int64)
        kits = _batch_kits(vote, kit_of_layout, len(kit_names), batch_size)
        b = 0
        while b < len(kits):                                    # consecutive batches that chose the same kit: one call
            e = b
            while e + 1 < len(kits) and kits[e + 1] == kits[b]:
                e += 1
            lo, hi = b * batch_size, min(n, (e + 1) * batch_size)
            plan.detect(win5[lo:hi], tail3[lo:hi], wlen[lo:hi], read_len[lo:hi], tables.kit_subset(kit_names[kits[b]]),
                        out=results[lo:hi])
            b = e + 1
    if getattr(scanner, "scan_middle_adapter", False):
        if chunk is None or qcat_config is None:
            raise ValueError("--detect-middle needs the chunk's records (the read bodies are scanned)")
        _middle_scan_chunk(scanner, chunk, results, qcat_config)
    if getattr(scanner, "enable_filter_barcodes", False) and not nobatch:
        _filter_barcodes(tables, results, batch_size)
    return results


def _filter_barcodes(tables, results, batch_size):
    """filter_barcodes per CLI batch (scanner_base.py:680-712): barcode ids seen in <= int(5 % of the most frequent
    key's count) reads -- "0" = unclassified counts as a key -- become empty results (trims reset to 0)."""
    n = len(results)
    called = results["barcode"] >= 0                            # simple mode: a barcode without an adapter (layout -1)
    key = np.maximum(results["layout"], 0).astype(np.int64) * (1 << 32) + results["barcode"].astype(np.int64)
    ids = np.zeros(n, dtype=np.int64)                           # 0 = unclassified (barcode ids start at 1)
    uniq, inverse = np.unique(key[called], return_inverse=True)
    id_of = {}
    codes = np.zeros(len(uniq), dtype=np.int64)
    for j, k in enumerate(uniq):
        b = tables.barcode_object(int(k >> 32), int(k & 0xffffffff))
        ident = (b[0].id, b[1].id) if tables.mode == 1 else b.id
        codes[j] = id_of.setdefault(ident, len(id_of) + 1)
    ids[called] = codes[inverse]
    for lo in range(0, n, batch_size):
        part = ids[lo:lo + batch_size]
        counts = np.bincount(part)
        min_count = int(counts.max() * 0.05)
        drop = (counts[part] <= min_count) & (part > 0)
        if drop.any():
            view = results[lo:lo + batch_size]
            view[drop] = np.array((-1, -1, 0.0, 0, 0, 0, 1), dtype=_ffi.RESULT_DTYPE)


class _BatchStraddler(object):
    """Batch semantics of the CLI loop (cli.py:500-513: consecutive batches of `batch_size` reads, each with its own kit
    vote / barcode filter) over chunks that end anywhere.  The windows of the reads behind the last complete batch are
    kept (`tail`) and scored together with the following chunk's; a chunk is handed on once all of its reads have
    results, so chunks leave in input order, at most one scoring call later than they arrived."""

    def __init__(self, batch_size):
        self.batch_size = int(batch_size)
        self.open = []            # [chunk, read_len, results, reads filled so far], in input order
        self.tail = None          # packed windows of the reads that do not have results yet

    def _pending(self, packed):
        if self.tail is None:
            return packed
        return tuple(np.concatenate([t, p]) for t, p in zip(self.tail, packed))

    def _distribute(self, scored):
        """Hand `scored` (results of the first reads without results, in input order) to the open chunks."""
        finished, pos = [], 0
        while self.open:
            entry = self.open[0]
            chunk, read_len, results, filled = entry
            take = min(len(results) - filled, len(scored) - pos)
            results[filled:filled + take] = scored[pos:pos + take]
            entry[3] = filled + take
            pos += take
            if entry[3] < len(results):
                break
            finished.append((chunk, read_len, results))         # (a chunk without reads is complete at once)
            self.open.pop(0)
        assert pos == len(scored), "scored more reads than are waiting"
        return finished

    def add(self, chunk, packed, score):
        self.open.append([chunk, packed[3], np.zeros(len(packed[2]), dtype=_ffi.RESULT_DTYPE), 0])
        pending = self._pending(packed)
        n_full = len(pending[2]) // self.batch_size * self.batch_size
        if n_full == 0:
            self.tail = pending
            return self._distribute(np.zeros(0, dtype=_ffi.RESULT_DTYPE))
        scored = score(tuple(a[:n_full] for a in pending))
        self.tail = tuple(np.ascontiguousarray(a[n_full:]) for a in pending) if n_full < len(pending[2]) else None
        return self._distribute(scored)

    def finish(self, score):
        """The file's last, possibly short, batch."""
        if self.tail is None or not len(self.tail[2]):
            finished = self._distribute(np.zeros(0, dtype=_ffi.RESULT_DTYPE))
            assert not self.open
            return finished
        scored = score(self.tail)
        self.tail = None
        finished = self._distribute(scored)
        assert not self.open
        return finished

    def release_all(self, failed):
        if failed:
            for entry in self.open:
                entry[0].release()
            self.open = []


class _Writer(object):
    """Formats and writes one chunk's records; keeps the running counts."""

    def __init__(self, plan, trim, min_read_length, out_dir, tsv, output, threads):
        self.lib = _ffi.load()
        self.tables = plan.tables
        self.labels = _Labels(plan.tables)
        self.trim, self.min_read_length = bool(trim), int(min_read_length)
        self.out_dir, self.tsv, self.output = out_dir, tsv, output
        self.threads = threads
        self.files = {}
        self.counts = {}
        self.kit_counts = {}
        self.total = self.skipped = 0
        self._own_output = None
        self._buf = np.empty(0, dtype=np.uint8)                  # formatting buffer, reused so its pages stay mapped
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
        if isinstance(output, (str, bytes, os.PathLike)):
            self._own_output = self.output = open(output, "wb")
        if tsv is not None:
            self._write(tsv, b"name\tlength\tbarcode\tscore\tkit\tadapter_end\tcomment\n")

    def _buffer(self, size):
        if self._buf.size < size:
            self._buf = np.empty(int(size) + int(size) // 4 + 4096, dtype=np.uint8)
        return self._buf

    @staticmethod
    def _write(fh, data):
        if isinstance(fh, io.TextIOBase):
            fh.write(bytes(data).decode("latin-1"))
        else:
            fh.write(data)

    def emit(self, chunk, read_len, results):
        n = len(results)
        labels = self.labels
        called = results["barcode"] >= 0                        # simple mode: a barcode without an adapter (layout -1)
        key = np.maximum(results["layout"], 0).astype(np.int64) * (1 << 32) + results["barcode"].astype(np.int64)
        name_label = np.full(n, labels.none, dtype=np.int32)
        id_label = np.full(n, labels.none, dtype=np.int32)
        kit_label = np.full(n, labels.none, dtype=np.int32)
        uniq, inverse = np.unique(key[called], return_inverse=True)
        if len(uniq):
            trip = np.array([labels.lookup(int(k)) for k in uniq], dtype=np.int32).reshape(-1, 3)
            name_label[called], id_label[called], kit_label[called] = trip[inverse, 0], trip[inverse, 1], trip[inverse, 2]
        # adapter histogram key (cli.py:369-377): the kit of the detected adapter, also for reads without a barcode
        with_adapter = results["layout"] >= 0
        strings = labels.strings

        if self.trim:
            a = np.clip(results["trim5p"].astype(np.int64), 0, read_len)
            b = np.maximum(np.clip(results["trim3p"].astype(np.int64), 0, read_len), a)
            out_len = b - a
        else:
            out_len = read_len
        kept = out_len >= self.min_read_length
        self.total += n
        self.skipped += int((~kept).sum())
        for index, cnt in zip(*np.unique(name_label[kept], return_counts=True)):
            self.counts[strings[index]] = self.counts.get(strings[index], 0) + int(cnt)
        kit_of_layout = [str(layout.kit) for layout in self.tables.layouts]
        for index, cnt in zip(*np.unique(results["layout"][kept & with_adapter], return_counts=True)):
            self.kit_counts[kit_of_layout[index]] = self.kit_counts.get(kit_of_layout[index], 0) + int(cnt)
        no_adapter = int((kept & ~with_adapter).sum())
        if no_adapter:
            self.kit_counts["none"] = self.kit_counts.get("none", 0) + no_adapter

        if not (self.tsv is not None or self.out_dir or self.output is not None):
            return
        lib = self.lib
        recs = np.ascontiguousarray(chunk.recs)
        results = np.ascontiguousarray(results)
        blob, off = labels.table()
        kept_u8 = np.zeros(n, dtype=np.uint8)
        need = ctypes.c_int64(0)
        if self.tsv is not None:
            tsv_label = np.where(called, id_label, -1).astype(np.int32)
            cap = int(recs["title_len"].sum()) + 160 * n + 64
            out = self._buffer(cap)
            _check_io(lib.qcb_format_tsv(_vp(chunk.data), _vp(recs), _vp(results), _vp(tsv_label), _vp(kit_label), n, _vp(blob),
                                         _vp(off), len(strings), 1 if self.trim else 0, self.min_read_length, _vp(out), cap,
                                         ctypes.byref(need), _vp(kept_u8), self.threads))
            self._write(self.tsv, out[:need.value].data)
        if self.out_dir:
            n_bins = len(strings)
            bin_bytes = np.zeros(n_bins, dtype=np.int64)
            bin_off = np.zeros(n_bins, dtype=np.int64)
            args = (_vp(chunk.data), _vp(recs), _vp(results), _vp(name_label), n, n_bins, 1 if chunk.fastq else 0,
                    1 if self.trim else 0, self.min_read_length, _vp(bin_bytes))
            _check_io(lib.qcb_format_records(*args, None, 0, _vp(bin_off), _vp(kept_u8), self.threads))
            out = self._buffer(int(bin_bytes.sum()) + 1)
            _check_io(lib.qcb_format_records(*args, _vp(out), int(out.size), _vp(bin_off), _vp(kept_u8), self.threads))
            fds = np.full(n_bins, -1, dtype=np.int32)
            for index in np.nonzero(bin_bytes)[0]:
                name = strings[index].replace("/", "_")
                if name not in self.files:
                    self.files[name] = os.open(os.path.join(self.out_dir, name + (".fastq" if chunk.fastq else ".fasta")),
                                               os.O_WRONLY | os.O_CREAT | os.O_TRUNC | os.O_APPEND, 0o666)
                fds[index] = self.files[name]
            _check_io(lib.qcb_write_bins(_vp(fds), _vp(out), _vp(bin_off), _vp(bin_bytes), n_bins, self.threads))
        elif self.output is not None:
            # cli.py:552: the single stream is written when there is no -b folder
            args = (_vp(chunk.data), _vp(recs), _vp(results), _vp(id_label), n, _vp(blob), _vp(off), len(strings),
                    1 if chunk.fastq else 0, 1 if self.trim else 0, self.min_read_length)
            _check_io(lib.qcb_format_stream(*args, None, 0, ctypes.byref(need), _vp(kept_u8), self.threads))
            out = self._buffer(need.value + 1)
            _check_io(lib.qcb_format_stream(*args, _vp(out), int(out.size), ctypes.byref(need), _vp(kept_u8), self.threads))
            self._write(self.output, out[:need.value].data)

    def close(self):
        for fd in self.files.values():
            os.close(fd)
        if self._own_output is not None:
            self._own_output.close()


_STOP = object()


def demux_file(path, scanner, qcat_config=None, batch_size=4000, trim=False, min_read_length=0, out_dir=None, tsv=None,
               output=None, nobatch=False, chunk_bytes=None, threads=None, keep_records=True):
    """Demultiplex a FASTQ / FASTA file like `qcat -f path [-b out_dir] [--tsv] [-o output] [--trim]` (cli.py:445-563);
    path "-" reads standard input, the CLI's default when no file is given.

    scanner: a qcat_b200 (or drop-in patched qcat) scanner.  Reads are scored in batches of `batch_size` (4000 in the
    CLI; `nobatch` = single-read mode without the kit vote).  out_dir: per-barcode files as with `-b`; tsv: a file
    object receiving the `--tsv` table; output: path or file object for the CLI's single output stream (records with
    `barcode=<id>` comments; written when out_dir is not given, like the CLI).  Returns {"reads", "skipped",
    "barcodes": {name: count}, "kits": {kit: count}, "records"} (records = all qcb_result rows unless
    keep_records=False).

    Three pipeline stages on three threads: (1) native read + parallel index + window packing of chunk k+1, (2) device
    scoring of chunk k, (3) native formatting + file writes of chunk k-1.
    """
    from qcat_b200 import scanner as qscanner
    qcat_config = qcat_config or qscanner._default_config()
    plan = scanner._plan_for(qcat_config)
    threads = _threads(threads)
    writer = _Writer(plan, trim, min_read_length, out_dir, tsv, output, threads)
    all_records = []
    packed_q = queue.Queue(maxsize=2)
    scored_q = queue.Queue(maxsize=2)
    failure = []
    # Batches only matter when there is a kit vote or a per-batch barcode filter.  The reader then still cuts the file
    # wherever records end: the windows (332 B per read) of the reads behind a chunk's last complete batch wait for the
    # next chunk and are scored with it (`_BatchStraddler`), so no file bytes are moved to keep batches aligned.  Only
    # --detect-middle, which needs the bytes of every read of a batch at scoring time, asks the reader for aligned chunks
    # (those carry up to one batch of bytes from chunk to chunk, so they are made larger).
    batched = not nobatch and ((plan.tables.mode != 2 and len(set(l.kit for l in scanner.layouts)) > 1) or
                               getattr(scanner, "enable_filter_barcodes", False))
    straddle = batched and not getattr(scanner, "scan_middle_adapter", False)
    # windows travel to the device in the plan's 4-bit form (two base classes per byte) when it has one
    multi_kit = plan.tables.mode != 2 and len(set(l.kit for l in scanner.layouts)) > 1
    four_bit = (hasattr(plan, "detect4") and plan.base_classes() is not None and
                (nobatch or not multi_kit or (plan.tables.kit_index()[1] is not None and hasattr(plan, "detect_auto4"))))
    classes = plan.base_classes() if four_bit else None
    multiple_of = batch_size if (batched and not straddle) else 1
    if chunk_bytes is None:
        chunk_bytes = (256 << 20) if multiple_of > 1 else (64 << 20)

    def produce():
        reader = None
        try:
            reader = Reader(path, chunk_bytes, threads)
            for chunk in reader.chunks(multiple_of):
                if failure:
                    chunk.release()
                    break
                packed = pack_windows(chunk.data, chunk.recs, qcat_config.max_align_length, threads, classes=classes)
                packed_q.put((chunk, packed))
        except BaseException as exc:                           # noqa: BLE001 -- re-raised on the caller's thread
            failure.append(exc)
        finally:
            packed_q.put(_STOP)
            done.wait()                                        # the reader owns the chunks: it must outlive all of them
            if reader is not None:
                reader.close()

    def consume():
        try:
            while True:
                item = scored_q.get()
                if item is _STOP:
                    return
                chunk, read_len, results = item
                try:
                    if not failure:
                        writer.emit(chunk, read_len, results)
                finally:
                    chunk.release()
        except BaseException as exc:                           # noqa: BLE001
            failure.append(exc)
            while scored_q.get() is not _STOP:                 # drain so the scoring thread never blocks
                pass

    done = threading.Event()
    producer = threading.Thread(target=produce, name="qcb-ingest", daemon=True)
    consumer = threading.Thread(target=consume, name="qcb-egress", daemon=True)
    producer.start()
    consumer.start()
    drained = False
    straddler = _BatchStraddler(batch_size) if straddle else None

    def emit(chunk, read_len, results):
        if keep_records:
            all_records.append(results)
        scored_q.put((chunk, read_len, results))

    try:
        while True:
            item = packed_q.get()
            if item is _STOP:
                drained = True
                if straddler is not None and not failure:
                    try:
                        for done_chunk in straddler.finish(lambda packed: _score_chunk(scanner, plan, packed, batch_size, nobatch, four_bit=four_bit)):
                            emit(*done_chunk)
                    except BaseException as exc:               # noqa: BLE001
                        failure.append(exc)
                break
            chunk, packed = item
            if failure:
                chunk.release()
                continue
            try:
                if straddler is not None:
                    finished = straddler.add(chunk, packed, lambda packed: _score_chunk(scanner, plan, packed, batch_size, nobatch, four_bit=four_bit))
                else:
                    finished = [(chunk, packed[3], _score_chunk(scanner, plan, packed, batch_size, nobatch, chunk, qcat_config, four_bit=four_bit))]
            except BaseException as exc:                       # noqa: BLE001
                failure.append(exc)
                chunk.release()
                continue
            for done_chunk in finished:
                emit(*done_chunk)
    finally:
        if straddler is not None:
            straddler.release_all(failed=bool(failure) or not drained)
        if not drained:                                        # interrupted while waiting: let the reader thread finish
            failure.append(sys.exc_info()[1] or RuntimeError("demux_file aborted"))
            while True:
                item = packed_q.get()
                if item is _STOP:
                    break
                item[0].release()
        scored_q.put(_STOP)
        consumer.join()
        done.set()
        producer.join()
        writer.close()
    if failure:
        raise failure[0]
    records = np.concatenate(all_records) if all_records else np.zeros(0, dtype=_ffi.RESULT_DTYPE)
    return {"reads": writer.total, "skipped": writer.skipped, "barcodes": writer.counts, "kits": writer.kit_counts,
            "records": records}
