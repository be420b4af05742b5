"""Test-side access to the CPU oracle (oracle/, TEST INFRASTRUCTURE ONLY) and the golden vectors."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build as oracle_build  # noqa: E402
sys.path.pop(0)

from qcat_b200 import _ffi, config, scanner  # noqa: E402
from qcat_b200.tables import Tables  # noqa: E402

RESULT_FIELDS = ("layout", "barcode", "barcode_score", "adapter_end", "trim5p", "trim3p", "exit_status")

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = ctypes.CDLL(oracle_build.build())
        for name in ("qo_sg", "qo_sg_stats", "qo_detect", "qo_scan", "qo_kit_vote", "qo_find_best_adapter_template"):
            getattr(lib, name).restype = None
        lib.qo_count_cells.restype = ctypes.c_int64
        _oracle = lib
    return _oracle


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def oracle_detect(tables, win5, tail3, wlen, read_len, subset=None, threads=None):
    """qo_detect (oracle/qcat_oracle.c) on packed windows -> structured array."""
    lib = oracle_lib()
    st, keep = _ffi.tables_struct(tables)          # qo_tables has the same field layout as qcb_tables
    n = int(len(wlen))
    out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    win5 = np.ascontiguousarray(win5, dtype=np.uint8)
    tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
    wlen = np.ascontiguousarray(wlen, dtype=np.int32)
    read_len = np.ascontiguousarray(read_len, dtype=np.int64)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    lib.qo_detect(ctypes.byref(st), _vp(win5), _vp(tail3), ctypes.c_int(win5.shape[1]), _vp(wlen), _vp(read_len),
                  ctypes.c_int64(n), _vp(sub), ctypes.c_int(0 if sub is None else sub.size), _vp(out),
                  ctypes.c_int(threads or os.cpu_count() or 1))
    return out


def oracle_scan(tables, windows, subset=None, threads=None):
    """qo_scan on a list of already oriented windows (str / bytes, any length) -> structured array."""
    lib = oracle_lib()
    st, keep = _ffi.tables_struct(tables)
    raw = [w if isinstance(w, bytes) else (w or "").encode("latin-1", "replace") for w in windows]
    n = len(raw)
    stride = max([len(r) for r in raw] + [1])
    buf = np.zeros((n, stride), dtype=np.uint8)
    wlen = np.zeros(n, dtype=np.int32)
    for i, r in enumerate(raw):
        wlen[i] = len(r)
        buf[i, :len(r)] = np.frombuffer(r, dtype=np.uint8)
    out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    lib.qo_scan(ctypes.byref(st), _vp(buf), ctypes.c_int64(stride), _vp(wlen), ctypes.c_int64(n), _vp(sub),
                ctypes.c_int(0 if sub is None else sub.size), _vp(out), ctypes.c_int(threads or os.cpu_count() or 1))
    return out


def oracle_kit_vote(tables, win5, tail3, wlen, threads=None):
    lib = oracle_lib()
    st, keep = _ffi.tables_struct(tables)
    n = int(len(wlen))
    vote = np.zeros(n, dtype=np.int32)
    win5 = np.ascontiguousarray(win5, dtype=np.uint8)
    tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
    wlen = np.ascontiguousarray(wlen, dtype=np.int32)
    lib.qo_kit_vote(ctypes.byref(st), _vp(win5), _vp(tail3), ctypes.c_int(win5.shape[1]), _vp(wlen), ctypes.c_int64(n),
                    _vp(vote), ctypes.c_int(threads or os.cpu_count() or 1))
    return vote


def oracle_detect_auto(tables, win5, tail3, wlen, read_len, batch_size=4000, threads=None, return_kits=False):
    """CPU oracle of the auto-kit flow (detect_barcode_batch, scanner_base.py:714-733): per batch of `batch_size` reads
    the kit vote over all layouts (qo_kit_vote + get_most_abundant_kits), then qo_detect restricted to that kit."""
    n = int(len(wlen))
    names = list(tables.kit_names)
    kit_names = list(dict.fromkeys(names))
    out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    kits = []
    vote = oracle_kit_vote(tables, win5, tail3, wlen, threads=threads)
    for lo in range(0, n, batch_size):
        hi = min(n, lo + batch_size)
        kit = kit_from_votes(vote[lo:hi], names)
        kits.append(kit_names.index(kit))
        out[lo:hi] = oracle_detect(tables, win5[lo:hi], tail3[lo:hi], wlen[lo:hi], read_len[lo:hi],
                                   subset=tables.kit_subset(kit), threads=threads)
    return (out, np.asarray(kits, dtype=np.int32)) if return_kits else out


def oracle_count_cells(tables, win5, tail3, wlen, subset=None):
    lib = oracle_lib()
    st, keep = _ffi.tables_struct(tables)
    n = int(len(wlen))
    sub = np.arange(tables.n_layouts, dtype=np.int32) if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    full = ctypes.c_int64(0)
    cells = lib.qo_count_cells(ctypes.byref(st), _vp(np.ascontiguousarray(win5)), _vp(np.ascontiguousarray(tail3)),
                               ctypes.c_int(win5.shape[1]), _vp(np.ascontiguousarray(wlen, dtype=np.int32)),
                               ctypes.c_int64(n), _vp(sub), ctypes.c_int(sub.size), ctypes.byref(full))
    return int(cells), int(full.value)


def oracle_sg(query, ref, open, extend, matrix):
    lib = oracle_lib()
    size, mat, mapper = config.matrix_arrays(matrix)
    q = query if isinstance(query, bytes) else query.encode("latin-1", "replace")
    r = ref if isinstance(ref, bytes) else ref.encode("latin-1", "replace")
    sc, eq, er = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    mat = np.ascontiguousarray(mat, dtype=np.int32)
    mapper = np.ascontiguousarray(mapper, dtype=np.uint8)
    lib.qo_sg(q, len(q), r, len(r), int(open), int(extend), _vp(mat), size, _vp(mapper),
              ctypes.byref(sc), ctypes.byref(eq), ctypes.byref(er))
    return sc.value, eq.value, er.value


def oracle_sg_stats(query, ref, open, extend, matrix):
    """qo_sg_stats -> (score, end_query, end_ref, matches, similar, length)."""
    lib = oracle_lib()
    size, mat, mapper = config.matrix_arrays(matrix)
    q = query if isinstance(query, bytes) else query.encode("latin-1", "replace")
    r = ref if isinstance(ref, bytes) else ref.encode("latin-1", "replace")
    out = [ctypes.c_int32() for _ in range(6)]
    mat = np.ascontiguousarray(mat, dtype=np.int32)
    mapper = np.ascontiguousarray(mapper, dtype=np.uint8)
    lib.qo_sg_stats(q, len(q), r, len(r), int(open), int(extend), _vp(mat), size, _vp(mapper), *[ctypes.byref(v) for v in out])
    return tuple(v.value for v in out)


def load_golden_simple():
    data = np.load(os.path.join(ROOT, "tests", "golden", "golden_simple_v1.npz"))
    return data, json.loads(bytes(data["cases"]).decode())["cases"]


def simple_tables_for_case(case):
    sc = scanner.BarcodeScannerSimple(min_quality=case["min_quality"], kit=case["kit"])
    assert len(sc.barcodes) == case["n_barcodes"]
    return Tables.simple(sc.barcodes, config.qcatConfig(), sc.min_quality), sc


def load_golden():
    data = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    meta = json.loads(bytes(data["cases"]).decode())
    return data, meta["cases"], meta["ranges"]


def scanner_for_case(case, device=None):
    cls = scanner.BarcodeScannerDual if case["mode"] == "dual" else scanner.BarcodeScannerEPI2ME
    sc = cls(min_quality=case["min_quality"], kit=case["kit"], device=device)
    assert [l.kit for l in sc.layouts] == case["layout_kits"], "layout order differs from the golden run"
    return sc


def tables_for_case(case):
    sc = scanner_for_case(case)
    return Tables(sc.layouts, config.qcatConfig(), case["mode"], sc.min_quality), sc


def kit_from_votes(vote, names):
    return scanner.GpuScannerMixin._kit_from_votes(vote, names)


def assert_records_equal(got, want, what=""):
    for f in RESULT_FIELDS:
        a, b = got[f], want[f]
        if f == "barcode_score":
            same = a.view(np.int64) == b.view(np.int64)          # bit-exact doubles
        else:
            same = a == b
        if not same.all():
            bad = np.nonzero(~same)[0]
            i = int(bad[0])
            raise AssertionError("%s: field %s differs on %d/%d records; first at %d: got %r want %r" %
                                 (what, f, bad.size, len(want), i, got[i], want[i]))


class OraclePlan(object):
    """The slice of engine.DevicePlan the host layers use (tables, detect, kit_vote), computed by the CPU oracle: lets the
    CPU suite drive qcat_b200.fastx.demux_file and the drop-in's Python glue without a GPU.  TESTS ONLY."""

    def __init__(self, tables):
        self.tables = tables
        self.calls = 0

    def detect(self, win5, tail3, wlen, read_len, subset=None, out=None):
        self.calls += 1
        got = oracle_detect(self.tables, win5, tail3, wlen, read_len, subset)
        if out is not None:
            out[...] = got
            return out
        return got

    def kit_vote(self, win5, tail3, wlen):
        return oracle_kit_vote(self.tables, win5, tail3, wlen)

    def scan_windows(self, windows, subset=None):
        self.calls += 1
        return oracle_scan(self.tables, windows, subset)

    def detect_auto(self, win5, tail3, wlen, read_len, kit_of_layout, batch_size, out=None, return_kits=False):
        self.calls += 1
        assert list(kit_of_layout) == list(self.tables.kit_index()[1])
        got, kits = oracle_detect_auto(self.tables, win5, tail3, wlen, read_len, batch_size=int(batch_size), return_kits=True)
        if out is not None:
            out[...] = got
            got = out
        return (got, kits) if return_kits else got


def random_kit(rng, layout_cls, barcode_cls, double=False):
    """Random kit geometry for fuzzing (a `random.Random`): 1-4 layouts with flanks of 0-45 nt, barcode placeholders of
    12-30 nt, 1-40 barcodes per set -- some sets sharing a prefix / suffix (the shared-context columns then reach into
    the barcodes), some layouts differing only in one flank base.  Built from the given AdapterLayout / Barcode
    classes, so the same kit can be made of the reference's objects and of qcat_b200's mirrors."""
    def seq(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    def barcode_set(length, count):
        share_head = seq(rng.choice([0, 0, 2, 5]))[:length - 1]
        share_tail = seq(rng.choice([0, 0, 3]))[:max(0, length - 1 - len(share_head))]
        out, seen = [], set()
        while len(out) < count:
            s = share_head + seq(length - len(share_head) - len(share_tail)) + share_tail
            if s not in seen:
                seen.add(s)
                out.append(barcode_cls("barcode%02d" % (len(out) + 1), len(out) + 1, s, True))
        return out

    layouts = []
    n_layouts = rng.randrange(1, 5)
    blen = rng.choice([12, 16, 20, 24, 24, 24, 27, 30])
    count = rng.choice([1, 2, 3, 7, 12, 24, 33, 40])
    set1 = barcode_set(blen, count)
    set2 = barcode_set(rng.choice([16, 24]), rng.choice([2, 5, 24])) if double else None
    base_left, base_mid, base_right = seq(rng.randrange(0, 46)), seq(rng.randrange(4, 25)), seq(rng.randrange(0, 46))
    for li in range(n_layouts):
        left, right = base_left, base_right
        if li and rng.random() < 0.5:                      # a sibling layout: one flank base changed
            if left:
                p = rng.randrange(len(left))
                left = left[:p] + rng.choice("ACGT") + left[p + 1:]
        elif li:
            left, right = seq(rng.randrange(0, 46)), seq(rng.randrange(0, 46))
        sequence = left + "N" * blen
        if double:
            sequence += base_mid + "N" * len(set2[0].sequence)
        sequence += right
        layouts.append(layout_cls("KIT%d" % li, sequence, set1 if li % 2 == 0 else barcode_set(blen, count), set2,
                                  "random kit %d" % li))
    return layouts
