"""Flatten a scanner (layouts + barcode sets + qcatConfig) into the read-only tables the device plan holds.

Works on the reference's own `AdapterLayout` / `Barcode` / `qcatConfig` objects as well as on the mirrors in
this package -- only the reference's accessor names are used.  The same flattened arrays feed the C-ABI
(`qcb_tables`, include/qcat_b200.h) and, in tests, the CPU oracle (`qo_tables`).
"""
import numpy as np

from qcat_b200.config import matrix_arrays

MODE_EPI2ME = 0
MODE_DUAL = 1
MODE_SIMPLE = 2

# utils.revcomp's translation table (reference utils.py:26-27): ACGT + IUPAC, both cases; anything else unchanged.
_COMP_FROM = b"ACGTacgtRYMKrymkVBHDvbhd"
_COMP_TO = b"TGCAtgcaYRKMyrkmBVDHbvdh"
COMPLEMENT = np.arange(256, dtype=np.uint8)
COMPLEMENT[np.frombuffer(_COMP_FROM, dtype=np.uint8)] = np.frombuffer(_COMP_TO, dtype=np.uint8)


def _ascii(seq):
    return seq.encode("latin-1", "replace")


class _SimpleLayout(object):
    """Placeholder layout of simple mode (scanner_simple.py): no adapter, barcode set 0 = the scanner's `barcodes`,
    aligned bare (no context) against the whole window."""
    kit = "simple"
    trim_offset = 0

    def __init__(self, barcodes):
        self._barcodes = list(barcodes)

    def get_adapter_sequences(self):
        return "N"

    def get_adapter_length(self):
        return 1

    def get_barcode_length(self, k):
        return 0

    def get_barcode_end(self, k):
        return -1

    def is_double_barcode(self):
        return False

    def get_barcode_set(self, k):
        return self._barcodes if k == 0 else None

    def get_upstream_context(self, n, k):
        return ""

    def get_downstream_context(self, n, k):
        return ""


class Tables(object):
    """Numpy arrays (see qcb_tables) plus the Python objects the result records index into."""

    @classmethod
    def simple(cls, barcodes, qcat_config, min_quality):
        """Tables of a simple-mode scanner (scanner_simple.py): one placeholder layout, group 0 = the bare barcodes."""
        if not barcodes:
            raise ValueError("simple mode needs at least one barcode")
        return cls([_SimpleLayout(barcodes)], qcat_config, MODE_SIMPLE, min_quality)

    def __init__(self, layouts, qcat_config, mode, min_quality, barcodes_override=None):
        self.layouts = list(layouts)
        self.mode = MODE_DUAL if mode in (MODE_DUAL, "dual") else (MODE_SIMPLE if mode in (MODE_SIMPLE, "simple") else MODE_EPI2ME)
        self.min_quality = float(min_quality)
        cfg = qcat_config
        self.max_align_length = int(cfg.max_align_length)
        self.barcode_extension = int(cfg.extracted_barcode_extension)
        self.adapter_open = int(cfg.gap_open)
        self.adapter_extend = int(cfg.gap_extend)
        self.barcode_open = 1          # hard-coded at scanner_base.py:115-116
        self.barcode_extend = 1
        self.amat_size, self.amat, self.amap = matrix_arrays(cfg.matrix)
        self.bmat_size, self.bmat, self.bmap = matrix_arrays(cfg.matrix_barcode)
        self.comp = COMPLEMENT.copy()
        ctx = int(cfg.barcode_context_length)

        n = len(self.layouts)
        self.n_layouts = n
        adapter_off = [0]
        adapter_seq = bytearray()
        self.denom = np.zeros(n, dtype=np.float64)
        self.bc_end = np.full(n * 2, -1, dtype=np.int32)
        self.bc_len = np.zeros(n * 2, dtype=np.int32)
        self.group = np.full(n * 2, -1, dtype=np.int32)
        self.trim_offset = np.zeros(n, dtype=np.int32)
        self.is_double = np.zeros(n, dtype=np.int32)
        self.kit_names = []

        group_off = [0]
        tmpl_off = [0]
        tmpl_seq = bytearray()
        tmpl_ident = []
        self.group_barcodes = []       # per group: the list of Barcode objects, in scoring order
        ident_class = {}               # Barcode.id -> small int (Python equality classes, :589)

        for i, layout in enumerate(self.layouts):
            seq = layout.get_adapter_sequences()
            adapter_seq += _ascii(seq)
            adapter_off.append(len(adapter_seq))
            bc_len = layout.get_barcode_length(0) + layout.get_barcode_length(1)
            a_len = layout.get_adapter_length()
            # get_norm_socre, scanner_base.py:308-310
            self.denom[i] = float((a_len - bc_len) * cfg.match + bc_len * cfg.nmatch)
            self.trim_offset[i] = int(layout.trim_offset)
            self.is_double[i] = 1 if layout.is_double_barcode() else 0
            self.kit_names.append(layout.kit)
            for k in (0, 1):
                self.bc_end[i * 2 + k] = layout.get_barcode_end(k)
                self.bc_len[i * 2 + k] = layout.get_barcode_length(k)
                barcode_set = layout.get_barcode_set(k)
                if barcode_set is None:
                    continue
                if barcodes_override:                      # scanner_epi2me.py:91-92
                    barcode_set = barcodes_override
                up = layout.get_upstream_context(ctx, k)
                down = layout.get_downstream_context(ctx, k)
                self.group[i * 2 + k] = len(self.group_barcodes)
                self.group_barcodes.append(list(barcode_set))
                for barcode in barcode_set:
                    tmpl_seq += _ascii(up + barcode.sequence + down)
                    tmpl_off.append(len(tmpl_seq))
                    tmpl_ident.append(ident_class.setdefault(barcode.id, len(ident_class)))
                group_off.append(len(tmpl_off) - 1)

        self.adapter_off = np.asarray(adapter_off, dtype=np.int32)
        self.adapter_seq = np.frombuffer(bytes(adapter_seq) or b"\0", dtype=np.uint8).copy()
        self.n_groups = len(self.group_barcodes)
        self.group_off = np.asarray(group_off, dtype=np.int32)
        self.tmpl_off = np.asarray(tmpl_off, dtype=np.int32)
        self.tmpl_seq = np.frombuffer(bytes(tmpl_seq) or b"\0", dtype=np.uint8).copy()
        self.tmpl_ident = np.asarray(tmpl_ident or [0], dtype=np.int32)
        self.n_templates = len(tmpl_off) - 1
        if len(ident_class) >= 65536:
            raise ValueError("too many distinct barcode ids")

    # ---- helpers used by the host API -------------------------------------------------------------

    def kit_subset(self, kit_name):
        """Indices of the layouts whose kit matches (BarcodeScanner.get_adapters, scanner_base.py:606-611)."""
        wanted = kit_name.lower()
        return [i for i, name in enumerate(self.kit_names) if name.lower() == wanted]

    def kit_index(self):
        """(kit names in first-seen order, int32 kit index per layout) -- the `kit_of_layout` table of qcb_detect_auto.
        Returns (names, None) when two kit names differ only by case: the reference counts votes per exact name but
        selects layouts case-insensitively (scanner_base.py:606-611), which one table cannot express."""
        names = list(dict.fromkeys(self.kit_names))
        if len(set(n.lower() for n in names)) != len(names):
            return names, None
        return names, np.array([names.index(k) for k in self.kit_names], dtype=np.int32)

    def group_size(self, layout_index, k):
        g = int(self.group[layout_index * 2 + k])
        return 0 if g < 0 else int(self.group_off[g + 1] - self.group_off[g])

    def barcode_object(self, layout_index, barcode_index):
        """Map a result record's (layout, barcode) back to the Barcode object(s) held by the layouts."""
        if self.mode == MODE_SIMPLE:              # no adapter: records carry layout -1
            return None if barcode_index < 0 else self.group_barcodes[0][barcode_index]
        if barcode_index < 0 or layout_index < 0:
            return None
        g1 = int(self.group[layout_index * 2])
        if self.mode == MODE_DUAL:
            g2 = int(self.group[layout_index * 2 + 1])
            n2 = int(self.group_off[g2 + 1] - self.group_off[g2])
            return self.group_barcodes[g1][barcode_index // n2], self.group_barcodes[g2][barcode_index % n2]
        return self.group_barcodes[g1][barcode_index]


def pack_windows(read_sequences, max_align_length, stride=None):
    """Cut reads into the two windows the scanner looks at (extract_align_sequence, scanner_base.py:223-244).

    Returns (win5, tail3, wlen, read_len, stride): win5[i] = read[:W], tail3[i] = read[-W:] (NOT reverse
    complemented -- the device does that), both in `stride`-byte slots (W rounded up to 16, so every slot is
    16-byte aligned for vector / bulk loads), wlen = min(len, W), read_len = len(read).
    """
    W = int(max_align_length)
    if stride is None:
        stride = max(16, (W + 15) // 16 * 16)
    if W <= 0:                     # length <= 0: the whole read is scanned (scanner_base.py:238)
        raise ValueError("max_align_length must be positive for the batched path")
    n = len(read_sequences)
    if n == 0:
        return (np.zeros((0, stride), dtype=np.uint8), np.zeros((0, stride), dtype=np.uint8),
                np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int64), stride)
    # One join per window side: slicing / padding stay inside CPython's string code (~1 us per read).
    seqs = [s if s else "" for s in read_sequences]
    read_len = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
    pad = "\0" * stride
    heads = "".join([(s[:W] + pad)[:stride] for s in seqs])
    tails = "".join([(s[-W:] + pad)[:stride] for s in seqs])
    win5 = np.frombuffer(_ascii(heads), dtype=np.uint8).reshape(n, stride).copy()
    tail3 = np.frombuffer(_ascii(tails), dtype=np.uint8).reshape(n, stride).copy()
    wlen = np.minimum(read_len, W).astype(np.int32)
    return win5, tail3, wlen, read_len, stride
