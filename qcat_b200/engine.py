"""Device plan: thin Python owner of a `qcb_plan` (include/qcat_b200.h).

Inputs and outputs are numpy arrays (host entry points) or raw device pointers (device entry points, e.g.
`tensor.data_ptr()` of torch tensors -- torch is only used by callers for device memory and streams).
"""
import ctypes
import os

import numpy as np

from qcat_b200 import _ffi
from qcat_b200.tables import Tables, pack_windows


def default_device():
    for key in ("QCAT_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


def device_count():
    return int(_ffi.load().qcb_device_count())


def _vp(arr):
    return ctypes.c_void_p(arr.ctypes.data)


def sg_batch(queries, refs, open, extend, matrix, device=None, stats=False):
    """Every query against every reference with parasail `sg` semantics -> (score, end_query, end_ref) int32
    arrays of shape (len(queries), len(refs)); stats=True (parasail `sg_stats`) adds (matches, similar, length).
    `matrix` is a ScoreMatrix or parasail-like Matrix."""
    from qcat_b200.config import matrix_arrays
    lib = _ffi.load()
    msize, mat, mapper = matrix_arrays(matrix)

    def concat(seqs):
        raw = [s if isinstance(s, bytes) else s.encode("latin-1", "replace") for s in seqs]
        off = np.zeros(len(raw) + 1, dtype=np.int32)
        if raw:
            off[1:] = np.cumsum([len(r) for r in raw])
        buf = np.frombuffer(b"".join(raw) or b"\0", dtype=np.uint8).copy()
        return buf, off

    qbuf, qoff = concat(queries)
    rbuf, roff = concat(refs)
    shape = (len(queries), len(refs))
    score = np.zeros(shape, dtype=np.int32)
    end_query = np.zeros(shape, dtype=np.int32)
    end_ref = np.zeros(shape, dtype=np.int32)
    mat = np.ascontiguousarray(mat, dtype=np.int32)
    mapper = np.ascontiguousarray(mapper, dtype=np.uint8)
    args = (default_device() if device is None else int(device), _vp(qbuf), _vp(qoff), len(queries), _vp(rbuf), _vp(roff),
            len(refs), int(open), int(extend), _vp(mat), msize, _vp(mapper), _vp(score), _vp(end_query), _vp(end_ref))
    if not stats:
        _ffi.check(lib.qcb_sg_batch(*args))
        return score, end_query, end_ref
    extra = [np.zeros(shape, dtype=np.int32) for _ in range(3)]
    _ffi.check(lib.qcb_sg_stats_batch(*args, *[_vp(a) for a in extra]))
    return (score, end_query, end_ref) + tuple(extra)


class DevicePlan(object):
    """Immutable device-side copy of a scanner's tables plus its workspace."""

    def __init__(self, tables, device=None):
        if not isinstance(tables, Tables):
            raise TypeError("tables must be a qcat_b200.tables.Tables")
        self._lib = _ffi.load()
        self.tables = tables
        self.device = default_device() if device is None else int(device)
        struct, self._keep = _ffi.tables_struct(tables)
        self._handle = self._lib.qcb_plan_create(ctypes.byref(struct), self.device)
        if not self._handle:
            raise _ffi.QcbError(_ffi.last_error() or "qcb_plan_create failed")

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.qcb_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = _ffi.QcbPlanInfo()
        _ffi.check(self._lib.qcb_plan_info(self._handle, ctypes.byref(out)))
        return {name: getattr(out, name) for name, _ in out._fields_}

    def set_force_generic(self, force):
        _ffi.check(self._lib.qcb_plan_set_force_generic(self._handle, int(force)))

    STAGES = ("orient", "adapter", "select", "barcode", "decide", "context")

    def set_profiling(self, enable):
        _ffi.check(self._lib.qcb_plan_set_profiling(self._handle, 1 if enable else 0))

    def stage_times(self, reset=True):
        """{stage: (milliseconds, kernel launches)} accumulated since the last reset (synchronises)."""
        ms = np.zeros(len(self.STAGES), dtype=np.float64)
        launches = np.zeros(len(self.STAGES), dtype=np.int64)
        _ffi.check(self._lib.qcb_plan_stage_times(self._handle, _vp(ms), _vp(launches), 1 if reset else 0))
        return {name: (float(ms[i]), int(launches[i])) for i, name in enumerate(self.STAGES)}

    @staticmethod
    def _subset(subset):
        if subset is None:
            return None, 0
        arr = np.ascontiguousarray(subset, dtype=np.int32)
        return arr, int(arr.size)

    # ---- host buffers ---------------------------------------------------------------------------

    def detect(self, win5, tail3, wlen, read_len, subset=None, out=None):
        """qcb_detect on host numpy arrays (see tables.pack_windows) -> structured array (RESULT_DTYPE)."""
        win5 = np.ascontiguousarray(win5, dtype=np.uint8)
        tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int64)
        n = int(wlen.shape[0])
        stride = int(win5.shape[1]) if win5.ndim == 2 else int(win5.size // max(n, 1))
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        sub, nsub = self._subset(subset)
        _ffi.check(self._lib.qcb_detect(self._handle, _vp(win5), _vp(tail3), stride, _vp(wlen), _vp(read_len), n,
                                        _vp(sub) if sub is not None else None, nsub, _vp(out)))
        return out

    def detect_auto(self, win5, tail3, wlen, read_len, kit_of_layout, batch_size, out=None, return_kits=False):
        """qcb_detect_auto on host numpy arrays: one pass over all layouts, per-batch kit vote on the device, detection
        restricted to every batch's kit.  kit_of_layout: Tables.kit_index()[1]."""
        win5 = np.ascontiguousarray(win5, dtype=np.uint8)
        tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int64)
        kits = np.ascontiguousarray(kit_of_layout, dtype=np.int32)
        if kits.size != self.tables.n_layouts:
            raise ValueError("kit_of_layout needs one entry per layout")
        n = int(wlen.shape[0])
        stride = int(win5.shape[1]) if win5.ndim == 2 else int(win5.size // max(n, 1))
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        batch_kit = np.zeros(max(1, (n + int(batch_size) - 1) // max(int(batch_size), 1)), dtype=np.int32)
        _ffi.check(self._lib.qcb_detect_auto(self._handle, _vp(win5), _vp(tail3), stride, _vp(wlen), _vp(read_len), n,
                                             _vp(kits), int(batch_size), _vp(out), _vp(batch_kit)))
        return (out, batch_kit) if return_kits else out

    # ---- 4-bit windows (two base classes per byte: half the host -> device traffic) --------------------

    def base_classes(self):
        """Byte -> class table (uint8[256]) of the 4-bit window format, or None when the plan's tables need more than
        16 classes (qcb_plan_base_classes)."""
        if not hasattr(self, "_classes"):
            cls = np.zeros(256, dtype=np.uint8)
            n = int(self._lib.qcb_plan_base_classes(self._handle, _vp(cls)))
            self._classes = cls if n > 0 else None
        return self._classes

    def pack4(self, windows, wlen, threads=0):
        """ASCII windows [n][stride] -> 4-bit windows [n][stride / 2] (qcb_pack_ascii4, host side)."""
        cls = self.base_classes()
        if cls is None:
            raise _ffi.QcbError("this plan has no 4-bit window format")
        windows = np.ascontiguousarray(windows, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        n, stride = windows.shape
        stride4 = (stride // 2 + 7) // 8 * 8
        out = np.zeros((n, stride4), dtype=np.uint8)
        rc = self._lib.qcb_pack_ascii4(_vp(windows), stride, _vp(wlen), n, _vp(cls), _vp(out), stride4, int(threads or os.cpu_count() or 1))
        if rc:
            raise _ffi.QcbError("qcb_pack_ascii4 failed")
        return out

    def detect4(self, win5p, tail3p, wlen, read_len, subset=None, out=None):
        """qcb_detect4: detect() on 4-bit windows (pack4 / fastx.pack_windows(..., classes=...))."""
        win5p = np.ascontiguousarray(win5p, dtype=np.uint8)
        tail3p = np.ascontiguousarray(tail3p, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int64)
        n = int(wlen.shape[0])
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        sub, nsub = self._subset(subset)
        _ffi.check(self._lib.qcb_detect4(self._handle, _vp(win5p), _vp(tail3p), int(win5p.shape[1]), _vp(wlen), _vp(read_len), n,
                                         _vp(sub) if sub is not None else None, nsub, _vp(out)))
        return out

    def detect_auto4(self, win5p, tail3p, wlen, read_len, kit_of_layout, batch_size, out=None, return_kits=False):
        win5p = np.ascontiguousarray(win5p, dtype=np.uint8)
        tail3p = np.ascontiguousarray(tail3p, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int64)
        kits = np.ascontiguousarray(kit_of_layout, dtype=np.int32)
        n = int(wlen.shape[0])
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        batch_kit = np.zeros(max(1, (n + int(batch_size) - 1) // max(int(batch_size), 1)), dtype=np.int32)
        _ffi.check(self._lib.qcb_detect_auto4(self._handle, _vp(win5p), _vp(tail3p), int(win5p.shape[1]), _vp(wlen), _vp(read_len),
                                              n, _vp(kits), int(batch_size), _vp(out), _vp(batch_kit)))
        return (out, batch_kit) if return_kits else out

    def detect4_device(self, d_win5p, d_tail3p, stride4, d_wlen, d_read_len, n_reads, d_out, subset=None, stream=0):
        sub, nsub = self._subset(subset)
        _ffi.check(self._lib.qcb_detect4_device(self._handle, ctypes.c_void_p(d_win5p), ctypes.c_void_p(d_tail3p), int(stride4),
                                                ctypes.c_void_p(d_wlen), ctypes.c_void_p(d_read_len), int(n_reads),
                                                _vp(sub) if sub is not None else None, nsub,
                                                ctypes.c_void_p(d_out), ctypes.c_void_p(stream)))

    def detect_reads(self, read_sequences, subset=None):
        win5, tail3, wlen, read_len, _ = pack_windows(read_sequences, self.tables.max_align_length)
        return self.detect(win5, tail3, wlen, read_len, subset)

    def scan_windows(self, windows, subset=None):
        """qcb_scan on a list of already-oriented window strings of any length -> structured array."""
        raw = [w if isinstance(w, bytes) else (w or "").encode("latin-1", "replace") for w in windows]
        n = len(raw)
        longest = max([len(r) for r in raw] + [1])
        stride = (longest + 15) // 16 * 16
        buf = np.zeros((n, stride), dtype=np.uint8)
        wlen = np.zeros(n, dtype=np.int32)
        for i, r in enumerate(raw):
            wlen[i] = len(r)
            buf[i, :len(r)] = np.frombuffer(r, dtype=np.uint8)
        out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        sub, nsub = self._subset(subset)
        _ffi.check(self._lib.qcb_scan(self._handle, _vp(buf), stride, _vp(wlen), n,
                                      _vp(sub) if sub is not None else None, nsub, _vp(out)))
        return out

    def kit_vote(self, win5, tail3, wlen):
        win5 = np.ascontiguousarray(win5, dtype=np.uint8)
        tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        n = int(wlen.shape[0])
        stride = int(win5.shape[1])
        vote = np.zeros(n, dtype=np.int32)
        _ffi.check(self._lib.qcb_kit_vote(self._handle, _vp(win5), _vp(tail3), stride, _vp(wlen), n, _vp(vote)))
        return vote

    # ---- device buffers (raw pointers) ----------------------------------------------------------

    def detect_device(self, d_win5, d_tail3, stride, d_wlen, d_read_len, n_reads, d_out, subset=None, stream=0):
        sub, nsub = self._subset(subset)
        _ffi.check(self._lib.qcb_detect_device(self._handle, ctypes.c_void_p(d_win5), ctypes.c_void_p(d_tail3), int(stride),
                                               ctypes.c_void_p(d_wlen), ctypes.c_void_p(d_read_len), int(n_reads),
                                               _vp(sub) if sub is not None else None, nsub,
                                               ctypes.c_void_p(d_out), ctypes.c_void_p(stream)))

    def detect_auto_device(self, d_win5, d_tail3, stride, d_wlen, d_read_len, n_reads, kit_of_layout, batch_size, d_out,
                           d_batch_kit=0, stream=0):
        kits = np.ascontiguousarray(kit_of_layout, dtype=np.int32)
        _ffi.check(self._lib.qcb_detect_auto_device(self._handle, ctypes.c_void_p(d_win5), ctypes.c_void_p(d_tail3), int(stride),
                                                    ctypes.c_void_p(d_wlen), ctypes.c_void_p(d_read_len), int(n_reads),
                                                    _vp(kits), int(batch_size), ctypes.c_void_p(d_out),
                                                    ctypes.c_void_p(d_batch_kit) if d_batch_kit else None, ctypes.c_void_p(stream)))

    def kit_vote_device(self, d_win5, d_tail3, stride, d_wlen, n_reads, d_vote, stream=0):
        _ffi.check(self._lib.qcb_kit_vote_device(self._handle, ctypes.c_void_p(d_win5), ctypes.c_void_p(d_tail3), int(stride),
                                                 ctypes.c_void_p(d_wlen), int(n_reads), ctypes.c_void_p(d_vote),
                                                 ctypes.c_void_p(stream)))

    def histogram_device(self, d_results, n_reads, layout_bin_base, d_counts, n_bins, stream=0):
        base = np.ascontiguousarray(layout_bin_base, dtype=np.int32)
        _ffi.check(self._lib.qcb_histogram_device(self._handle, ctypes.c_void_p(d_results), int(n_reads), _vp(base),
                                                  ctypes.c_void_p(d_counts), int(n_bins), ctypes.c_void_p(stream)))

    def histogram_layout(self):
        """(layout_bin_base, n_bins): bin 0 = unclassified, then one bin per (layout, barcode index)."""
        t = self.tables
        base = np.zeros(t.n_layouts, dtype=np.int32)
        total = 0
        for i in range(t.n_layouts):
            base[i] = total
            size = t.group_size(i, 0)               # simple mode: one placeholder layout, group 0 = the barcodes
            if t.mode == 1:
                size *= t.group_size(i, 1)
            total += size
        return base, total + 1


class MultiDevicePlan(object):
    """One DevicePlan per device of this process behind the DevicePlan interface (qcb_detect_multi /
    qcb_detect_auto_multi): reads are dealt to the devices in blocks, round-robin, every device runs its own copy /
    compute pipeline on its own host thread, and records land at their reads' positions.  `devices`: CUDA device
    indices; an index may repeat (two plans sharing one device -- how the sharding is tested on a one-GPU box)."""

    def __init__(self, tables, devices):
        devices = [int(d) for d in devices]
        if not devices:
            raise ValueError("MultiDevicePlan needs at least one device")
        self.tables = tables
        self.devices = devices
        self.plans = [DevicePlan(tables, device=d) for d in devices]
        self._lib = self.plans[0]._lib
        self._handles = (ctypes.c_void_p * len(self.plans))(*[p._handle for p in self.plans])
        self.device = devices[0]

    def close(self):
        for plan in self.plans:
            plan.close()

    def info(self):
        infos = [p.info() for p in self.plans]
        out = dict(infos[0])
        out["kernel_launches"] = sum(i["kernel_launches"] for i in infos)
        out["workspace_bytes"] = sum(i["workspace_bytes"] for i in infos)
        out["devices"] = list(self.devices)
        out["kernel_launches_per_device"] = [i["kernel_launches"] for i in infos]
        return out

    def set_force_generic(self, force):
        for plan in self.plans:
            plan.set_force_generic(force)

    @staticmethod
    def _host_arrays(win5, tail3, wlen, read_len):
        win5 = np.ascontiguousarray(win5, dtype=np.uint8)
        tail3 = np.ascontiguousarray(tail3, dtype=np.uint8)
        wlen = np.ascontiguousarray(wlen, dtype=np.int32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int64)
        n = int(wlen.shape[0])
        stride = int(win5.shape[1]) if win5.ndim == 2 else int(win5.size // max(n, 1))
        return win5, tail3, wlen, read_len, n, stride

    def detect(self, win5, tail3, wlen, read_len, subset=None, out=None):
        win5, tail3, wlen, read_len, n, stride = self._host_arrays(win5, tail3, wlen, read_len)
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        sub, nsub = DevicePlan._subset(subset)
        _ffi.check(self._lib.qcb_detect_multi(self._handles, len(self.plans), _vp(win5), _vp(tail3), stride, _vp(wlen),
                                              _vp(read_len), n, _vp(sub) if sub is not None else None, nsub, _vp(out)))
        return out

    def detect_auto(self, win5, tail3, wlen, read_len, kit_of_layout, batch_size, out=None, return_kits=False):
        win5, tail3, wlen, read_len, n, stride = self._host_arrays(win5, tail3, wlen, read_len)
        kits = np.ascontiguousarray(kit_of_layout, dtype=np.int32)
        if kits.size != self.tables.n_layouts:
            raise ValueError("kit_of_layout needs one entry per layout")
        if out is None:
            out = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
        batch_kit = np.zeros(max(1, (n + int(batch_size) - 1) // max(int(batch_size), 1)), dtype=np.int32)
        _ffi.check(self._lib.qcb_detect_auto_multi(self._handles, len(self.plans), _vp(win5), _vp(tail3), stride, _vp(wlen),
                                                   _vp(read_len), n, _vp(kits), int(batch_size), _vp(out), _vp(batch_kit)))
        return (out, batch_kit) if return_kits else out

    # calls that are not worth spreading (single windows, the fallback vote) run on the first device
    def scan_windows(self, windows, subset=None):
        return self.plans[0].scan_windows(windows, subset)

    def kit_vote(self, win5, tail3, wlen):
        return self.plans[0].kit_vote(win5, tail3, wlen)

    def histogram_layout(self):
        return self.plans[0].histogram_layout()

    def hist_allgather(self, d_counts, n_bins, d_gathered):
        """qcb_hist_allgather: d_counts[d] / d_gathered[d] are device pointers (ints) on device d."""
        counts = (ctypes.c_void_p * len(self.plans))(*[int(c) for c in d_counts])
        gathered = (ctypes.c_void_p * len(self.plans))(*[int(g) for g in d_gathered])
        _ffi.check(self._lib.qcb_hist_allgather(self._handles, len(self.plans), counts, int(n_bins), gathered))


def make_plan(tables, device=None):
    """DevicePlan for one device index (or None = default), MultiDevicePlan for a list / tuple of indices or "all"."""
    if isinstance(device, str) and device == "all":
        device = list(range(device_count()))
    if isinstance(device, (list, tuple)):
        return MultiDevicePlan(tables, device) if len(device) != 1 else DevicePlan(tables, device=device[0])
    return DevicePlan(tables, device=device)


def microbench_cell_rate(device=None):
    """(packed DP cell updates per second the SMs can issue, effective SM MHz) on this device."""
    cells = ctypes.c_double()
    mhz = ctypes.c_double()
    _ffi.check(_ffi.load().qcb_microbench_cell_rate(default_device() if device is None else int(device),
                                                    ctypes.byref(cells), ctypes.byref(mhz)))
    return cells.value, mhz.value
