"""ctypes binding of libqcat_b200.so (include/qcat_b200.h).  No CPU fallback: a missing library or a box
without a CUDA device raises instead of silently computing somewhere else."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QCAT_B200_LIB: another build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("QCAT_B200_LIB") or os.path.join(_HERE, "libqcat_b200.so")

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)


class QcbTables(ctypes.Structure):
    """qcb_tables (the CPU oracle's qo_tables has the same field layout)."""
    _fields_ = [
        ("max_align_length", ctypes.c_int32), ("barcode_extension", ctypes.c_int32),
        ("adapter_open", ctypes.c_int32), ("adapter_extend", ctypes.c_int32),
        ("barcode_open", ctypes.c_int32), ("barcode_extend", ctypes.c_int32),
        ("amat_size", ctypes.c_int32), ("amat", c_i32p), ("amap", c_u8p),
        ("bmat_size", ctypes.c_int32), ("bmat", c_i32p), ("bmap", c_u8p),
        ("comp", c_u8p),
        ("mode", ctypes.c_int32), ("min_quality", ctypes.c_double),
        ("n_layouts", ctypes.c_int32), ("adapter_off", c_i32p), ("adapter_seq", c_u8p),
        ("denom", c_f64p), ("bc_end", c_i32p), ("bc_len", c_i32p), ("group", c_i32p),
        ("trim_offset", c_i32p), ("is_double", c_i32p),
        ("n_groups", ctypes.c_int32), ("group_off", c_i32p), ("tmpl_off", c_i32p), ("tmpl_seq", c_u8p),
        ("tmpl_ident", c_i32p),
    ]


class QcbResult(ctypes.Structure):
    _fields_ = [("layout", ctypes.c_int32), ("barcode", ctypes.c_int32), ("barcode_score", ctypes.c_double),
                ("adapter_end", ctypes.c_int32), ("trim5p", ctypes.c_int32), ("trim3p", ctypes.c_int32),
                ("exit_status", ctypes.c_int32)]


RESULT_DTYPE = np.dtype([("layout", "<i4"), ("barcode", "<i4"), ("barcode_score", "<f8"), ("adapter_end", "<i4"),
                         ("trim5p", "<i4"), ("trim3p", "<i4"), ("exit_status", "<i4")], align=True)
assert RESULT_DTYPE.itemsize == ctypes.sizeof(QcbResult) == 32


RECORD_DTYPE = np.dtype([("title_off", "<i8"), ("title_len", "<i8"), ("seq_off", "<i8"), ("seq_span", "<i8"),
                         ("seq_len", "<i8"), ("qual_off", "<i8"), ("qual_span", "<i8")])      # qcb_fastx_record


class QcbPlanInfo(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("sm_count", ctypes.c_int32), ("fast_adapter", ctypes.c_int32),
                ("fast_barcode", ctypes.c_int32), ("max_group_size", ctypes.c_int32), ("n_templates", ctypes.c_int32),
                ("workspace_bytes", ctypes.c_int64), ("kernel_launches", ctypes.c_int64)]


def _ptr(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def tables_struct(tables):
    """Fill a QcbTables from a qcat_b200.tables.Tables; returns (struct, keepalive list)."""
    t = tables
    keep = []

    def arr(a, dtype, ctype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return _ptr(a, ctype)

    s = QcbTables()
    s.max_align_length = t.max_align_length
    s.barcode_extension = t.barcode_extension
    s.adapter_open, s.adapter_extend = t.adapter_open, t.adapter_extend
    s.barcode_open, s.barcode_extend = t.barcode_open, t.barcode_extend
    s.amat_size = t.amat_size
    s.amat = arr(t.amat, np.int32, ctypes.c_int32)
    s.amap = arr(t.amap, np.uint8, ctypes.c_uint8)
    s.bmat_size = t.bmat_size
    s.bmat = arr(t.bmat, np.int32, ctypes.c_int32)
    s.bmap = arr(t.bmap, np.uint8, ctypes.c_uint8)
    s.comp = arr(t.comp, np.uint8, ctypes.c_uint8)
    s.mode = t.mode
    s.min_quality = t.min_quality
    s.n_layouts = t.n_layouts
    s.adapter_off = arr(t.adapter_off, np.int32, ctypes.c_int32)
    s.adapter_seq = arr(t.adapter_seq, np.uint8, ctypes.c_uint8)
    s.denom = arr(t.denom, np.float64, ctypes.c_double)
    s.bc_end = arr(t.bc_end, np.int32, ctypes.c_int32)
    s.bc_len = arr(t.bc_len, np.int32, ctypes.c_int32)
    s.group = arr(t.group, np.int32, ctypes.c_int32)
    s.trim_offset = arr(t.trim_offset, np.int32, ctypes.c_int32)
    s.is_double = arr(t.is_double, np.int32, ctypes.c_int32)
    s.n_groups = t.n_groups
    s.group_off = arr(t.group_off, np.int32, ctypes.c_int32)
    s.tmpl_off = arr(t.tmpl_off, np.int32, ctypes.c_int32)
    s.tmpl_seq = arr(t.tmpl_seq, np.uint8, ctypes.c_uint8)
    s.tmpl_ident = arr(t.tmpl_ident, np.int32, ctypes.c_int32)
    return s, keep


# Every symbol include/qcat_b200.h declares (tests check the library exports all of them).
EXPORTS = ("qcb_device_count", "qcb_last_error", "qcb_version", "qcb_plan_create", "qcb_plan_destroy", "qcb_plan_info",
           "qcb_plan_set_force_generic", "qcb_plan_set_profiling", "qcb_plan_stage_times", "qcb_sg_batch", "qcb_sg_stats_batch", "qcb_scan", "qcb_detect", "qcb_detect_device", "qcb_detect_auto", "qcb_detect_auto_device", "qcb_plan_base_classes", "qcb_detect4", "qcb_detect4_device",
           "qcb_detect_auto4", "qcb_pack_ascii4", "qcb_pack_windows4", "qcb_detect_multi", "qcb_detect_auto_multi", "qcb_hist_allgather", "qcb_kit_vote",
           "qcb_kit_vote_device", "qcb_histogram_device", "qcb_microbench_cell_rate", "qcb_io_last_error", "qcb_fastx_index",
           "qcb_pack_windows", "qcb_format_records", "qcb_fastx_index_mt", "qcb_format_stream", "qcb_format_tsv",
           "qcb_write_bins", "qcb_reader_open", "qcb_reader_next", "qcb_chunk_data", "qcb_chunk_records", "qcb_chunk_release", "qcb_reader_close")

_lib = None


def load():
    """Load libqcat_b200.so (building is the job of __graft_entry__.build() / qcat_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libqcat_b200.so is missing at %s -- run `python qcat_b200/build.py`; "
                           "there is no CPU fallback for the CUDA path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.qcb_device_count.restype = ctypes.c_int
    lib.qcb_device_count.argtypes = []
    lib.qcb_last_error.restype = ctypes.c_char_p
    lib.qcb_last_error.argtypes = []
    lib.qcb_version.restype = ctypes.c_char_p
    lib.qcb_version.argtypes = []
    lib.qcb_plan_create.restype = vp
    lib.qcb_plan_create.argtypes = [ctypes.POINTER(QcbTables), ctypes.c_int]
    lib.qcb_plan_destroy.restype = None
    lib.qcb_plan_destroy.argtypes = [vp]
    lib.qcb_plan_info.restype = ctypes.c_int
    lib.qcb_plan_info.argtypes = [vp, ctypes.POINTER(QcbPlanInfo)]
    lib.qcb_plan_set_force_generic.restype = ctypes.c_int
    lib.qcb_plan_set_force_generic.argtypes = [vp, ctypes.c_int]
    lib.qcb_plan_set_profiling.restype = ctypes.c_int
    lib.qcb_plan_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.qcb_plan_stage_times.restype = ctypes.c_int
    lib.qcb_plan_stage_times.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.qcb_sg_batch.restype = ctypes.c_int
    lib.qcb_sg_batch.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_int32, vp, ctypes.c_int32, vp, vp, vp, vp]
    lib.qcb_sg_stats_batch.restype = ctypes.c_int
    lib.qcb_sg_stats_batch.argtypes = lib.qcb_sg_batch.argtypes + [vp, vp, vp]
    lib.qcb_detect.restype = ctypes.c_int
    lib.qcb_detect.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp]
    lib.qcb_scan.restype = ctypes.c_int
    lib.qcb_scan.argtypes = [vp, vp, ctypes.c_int32, vp, ctypes.c_int64, vp, ctypes.c_int32, vp]
    lib.qcb_detect_device.restype = ctypes.c_int
    lib.qcb_detect_device.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp, vp]
    lib.qcb_detect_auto.restype = ctypes.c_int
    lib.qcb_detect_auto.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp, vp]
    lib.qcb_detect_auto_device.restype = ctypes.c_int
    lib.qcb_detect_auto_device.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp, vp, vp]
    lib.qcb_plan_base_classes.restype = ctypes.c_int
    lib.qcb_plan_base_classes.argtypes = [vp, vp]
    lib.qcb_detect4.restype = ctypes.c_int
    lib.qcb_detect4.argtypes = lib.qcb_detect.argtypes
    lib.qcb_detect4_device.restype = ctypes.c_int
    lib.qcb_detect4_device.argtypes = lib.qcb_detect_device.argtypes
    lib.qcb_detect_auto4.restype = ctypes.c_int
    lib.qcb_detect_auto4.argtypes = lib.qcb_detect_auto.argtypes
    lib.qcb_pack_ascii4.restype = ctypes.c_int
    lib.qcb_pack_ascii4.argtypes = [vp, ctypes.c_int32, vp, ctypes.c_int64, vp, vp, ctypes.c_int32, ctypes.c_int32]
    lib.qcb_pack_windows4.restype = ctypes.c_int
    lib.qcb_pack_windows4.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp, vp, ctypes.c_int32]
    lib.qcb_detect_multi.restype = ctypes.c_int
    lib.qcb_detect_multi.argtypes = [vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp]
    lib.qcb_detect_auto_multi.restype = ctypes.c_int
    lib.qcb_detect_auto_multi.argtypes = [vp, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp, vp]
    lib.qcb_hist_allgather.restype = ctypes.c_int
    lib.qcb_hist_allgather.argtypes = [vp, ctypes.c_int32, vp, ctypes.c_int32, vp]
    lib.qcb_kit_vote.restype = ctypes.c_int
    lib.qcb_kit_vote.argtypes = [vp, vp, vp, ctypes.c_int32, vp, ctypes.c_int64, vp]
    lib.qcb_kit_vote_device.restype = ctypes.c_int
    lib.qcb_kit_vote_device.argtypes = [vp, vp, vp, ctypes.c_int32, vp, ctypes.c_int64, vp, vp]
    lib.qcb_histogram_device.restype = ctypes.c_int
    lib.qcb_histogram_device.argtypes = [vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int32, vp]
    lib.qcb_microbench_cell_rate.restype = ctypes.c_int
    lib.qcb_microbench_cell_rate.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.qcb_io_last_error.restype = ctypes.c_char_p
    lib.qcb_io_last_error.argtypes = []
    lib.qcb_fastx_index.restype = ctypes.c_int
    lib.qcb_fastx_index.argtypes = [vp, ctypes.c_int64, ctypes.c_int32, vp, ctypes.c_int64, vp, vp, vp]
    lib.qcb_pack_windows.restype = ctypes.c_int
    lib.qcb_pack_windows.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int32]
    lib.qcb_format_records.restype = ctypes.c_int
    lib.qcb_format_records.argtypes = [vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_int64, vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int32]
    i32, i64 = ctypes.c_int32, ctypes.c_int64
    lib.qcb_fastx_index_mt.restype = ctypes.c_int
    lib.qcb_fastx_index_mt.argtypes = [vp, i64, i32, vp, i64, vp, vp, vp, i32]
    lib.qcb_format_stream.restype = ctypes.c_int
    lib.qcb_format_stream.argtypes = [vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i64, vp, i64, vp, vp, i32]
    lib.qcb_format_tsv.restype = ctypes.c_int
    lib.qcb_format_tsv.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i64, vp, i64, vp, vp, i32]
    lib.qcb_write_bins.restype = ctypes.c_int
    lib.qcb_write_bins.argtypes = [vp, vp, vp, vp, i32, i32]
    lib.qcb_reader_open.restype = vp
    lib.qcb_reader_open.argtypes = [ctypes.c_char_p, i64, i32]
    lib.qcb_reader_next.restype = ctypes.c_int
    lib.qcb_reader_next.argtypes = [vp, i64, ctypes.POINTER(vp)]
    lib.qcb_chunk_data.restype = vp
    lib.qcb_chunk_data.argtypes = [vp, ctypes.POINTER(i64)]
    lib.qcb_chunk_records.restype = vp
    lib.qcb_chunk_records.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i32)]
    lib.qcb_chunk_release.restype = None
    lib.qcb_chunk_release.argtypes = [vp]
    lib.qcb_reader_close.restype = None
    lib.qcb_reader_close.argtypes = [vp]
    _lib = lib
    return lib


def last_error():
    msg = load().qcb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


class QcbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise QcbError(last_error() or "libqcat_b200 call failed (rc=%d)" % rc)
