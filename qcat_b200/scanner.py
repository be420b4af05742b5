"""Host-side scanner API: the reference's `BarcodeScanner` surface over the CUDA plan.

Mirrors qcat/scanner.py (factory, get_kits, get_modes), qcat/scanner_base.py:410-733 (BarcodeScanner),
qcat/scanner_epi2me.py and qcat/scanner_dual.py -- same class / method names, argument meaning and result
dicts -- so code written against qcat (cli.py:477-513, the reference tests) runs against it unchanged.
All alignment and per-read decision logic executes in libqcat_b200.so on the GPU; this module only packs
strings into window buffers and turns result records back into the reference's dict / Barcode / AdapterLayout
objects.  `GpuScannerMixin` holds every GPU-backed method and only relies on the attributes the reference's own
classes have, so qcat_b200.dropin can graft it onto an installed qcat.
"""
import logging
import operator
import os


from qcat_b200 import adapters
from qcat_b200 import config
from qcat_b200.adapters import Barcode
from qcat_b200.tables import Tables, pack_windows, COMPLEMENT

_COMP_TABLE = bytes(COMPLEMENT.tobytes())


def revcomp(seq):
    """qcat.utils.revcomp (reference utils.py:26-27)."""
    return seq.encode("latin-1", "replace").translate(_COMP_TABLE)[::-1].decode("latin-1")


def build_return_dict(best_barcode, best_barcode_score, best_adapter, best_adapter_end, exit_status,
                      trim5p=0, trim3p=0):
    """Result dict of the reference (scanner_base.py:362-390)."""
    return {"barcode": best_barcode,
            "barcode_score": best_barcode_score,
            "adapter": best_adapter,
            "adapter_end": best_adapter_end,
            "trim5p": trim5p,
            "trim3p": trim3p,
            "exit_status": exit_status}


def empty_return_dict():
    return build_return_dict(None, 0.0, None, 0, 1, trim5p=0, trim3p=0)


def _config_key(cfg):
    size_a, mat_a, map_a = config.matrix_arrays(cfg.matrix)
    size_b, mat_b, map_b = config.matrix_arrays(cfg.matrix_barcode)
    return (cfg.match, cfg.nmatch, cfg.mismatch, cfg.gap_open, cfg.gap_extend, cfg.max_align_length,
            cfg.extracted_barcode_extension, cfg.barcode_context_length,
            size_a, mat_a.tobytes(), map_a.tobytes(), size_b, mat_b.tobytes(), map_b.tobytes())


def _device_key(device):
    return tuple(device) if isinstance(device, (list, tuple)) else device


class GpuScannerMixin(object):
    """GPU-backed detect_barcode / detect_barcode_batch / scan for a reference-shaped scanner object
    (attributes used: layouts, min_quality, override_kit_name, barcodes, enable_filter_barcodes,
    scan_middle_adapter, get_name())."""

    device = None            # CUDA device index, a list of indices or "all" (one plan per device, reads sharded between
                             # them inside this process); None -> QCAT_B200_DEVICE / LOCAL_RANK / 0

    # ---- plan management ------------------------------------------------------------------------

    def _mode_name(self):
        name = self.get_name()
        return name if name in ("dual", "simple") else "epi2me"

    def _tables_for(self, qcat_config, layouts=None):
        """(cache key, Tables factory) of this scanner for a config: simple mode flattens `self.barcodes` only
        (scanner_simple.py never looks at a layout), the others their layouts (+ the epi2me barcode override)."""
        mode = self._mode_name()
        override = getattr(self, "barcodes", None)
        if mode == "simple":
            key = ("simple", tuple(override or ()), _config_key(qcat_config), float(self.min_quality), _device_key(self.device))
            return key, lambda: Tables.simple(override, qcat_config, self.min_quality)
        layouts = self.layouts if layouts is None else layouts
        key = (tuple(id(l) for l in layouts), _config_key(qcat_config), float(self.min_quality),
               None if not override else tuple(override), mode, _device_key(self.device))
        return key, lambda: Tables(layouts, qcat_config, mode, self.min_quality, override)

    def _plan_for(self, qcat_config, layouts=None):
        from qcat_b200.engine import make_plan
        key, make_tables = self._tables_for(qcat_config, layouts)
        cache = self.__dict__.setdefault("_qcb_plans", {})
        plan = cache.get(key)
        if plan is None:
            if len(cache) >= 4:
                cache.pop(next(iter(cache))).close()
            plan = cache[key] = make_plan(make_tables(), device=self.device)
        return plan

    def _subset_for(self, plan, kits):
        """Layout indices (in plan order) of `kits`, a list of layout objects."""
        index = {id(l): i for i, l in enumerate(plan.tables.layouts)}
        return [index[id(l)] for l in kits]

    # ---- record -> reference dict ---------------------------------------------------------------

    def _record_to_dict(self, plan, rec):
        layout_index = int(rec["layout"])
        if plan.tables.mode == 2:                 # simple: adapter None (scanner_simple.py:84-90)
            result = build_return_dict(plan.tables.barcode_object(-1, int(rec["barcode"])), float(rec["barcode_score"]), None,
                                       int(rec["adapter_end"]), int(rec["exit_status"]))
        elif layout_index < 0:
            result = empty_return_dict()
            result["exit_status"] = int(rec["exit_status"])
        else:
            barcode = plan.tables.barcode_object(layout_index, int(rec["barcode"]))
            if plan.tables.mode == 1:           # dual: synthesised pair (scanner_dual.py:132-136)
                first, second = barcode
                barcode = Barcode("barcode{:02d}/{:02d}".format(first.id, second.id),
                                  "{}/{}".format(first.id, second.id), None, True)
            result = build_return_dict(barcode, float(rec["barcode_score"]), plan.tables.layouts[layout_index],
                                       int(rec["adapter_end"]), int(rec["exit_status"]))
        result["trim5p"] = int(rec["trim5p"])
        result["trim3p"] = int(rec["trim3p"])
        return result

    def _records_to_dicts(self, plan, recs):
        """_record_to_dict for a whole batch: columns are converted to Python objects once (a structured-array row
        access costs ~10 us, this loop ~0.5 us per read)."""
        tables = plan.tables
        layouts = tables.layouts
        dual = tables.mode == 1
        if tables.mode == 2:
            return [self._record_to_dict(plan, rec) for rec in recs]
        barcodes = {}
        out = []
        for layout_index, barcode_index, score, adapter_end, trim5p, trim3p, exit_status in zip(
                recs["layout"].tolist(), recs["barcode"].tolist(), recs["barcode_score"].tolist(), recs["adapter_end"].tolist(),
                recs["trim5p"].tolist(), recs["trim3p"].tolist(), recs["exit_status"].tolist()):
            if layout_index < 0:
                out.append({"barcode": None, "barcode_score": 0.0, "adapter": None, "adapter_end": 0,
                            "trim5p": trim5p, "trim3p": trim3p, "exit_status": exit_status})
                continue
            key = (layout_index, barcode_index)
            barcode = barcodes.get(key)
            if barcode is None:
                barcode = tables.barcode_object(layout_index, barcode_index)
                if dual:                          # synthesised pair (scanner_dual.py:132-136)
                    first, second = barcode
                    barcode = Barcode("barcode{:02d}/{:02d}".format(first.id, second.id),
                                      "{}/{}".format(first.id, second.id), None, True)
                barcodes[key] = barcode
            out.append({"barcode": barcode, "barcode_score": score, "adapter": layouts[layout_index],
                        "adapter_end": adapter_end, "trim5p": trim5p, "trim3p": trim3p, "exit_status": exit_status})
        return out

    # ---- reference API ---------------------------------------------------------------------------

    def scan(self, read_sequence, read_qualities, bc_adapter_templates, nobc_adapter_templates,
             qcat_config=None):
        """One window (any length) against the given layouts (scanner_epi2me.py:33 / scanner_dual.py:35)."""
        qcat_config = qcat_config or _default_config()
        if self._mode_name() == "simple":             # scanner_simple.py:47-92 never looks at the templates
            plan = self._plan_for(qcat_config)
            return self._record_to_dict(plan, plan.scan_windows([read_sequence or ""], None)[0])
        if not isinstance(bc_adapter_templates, list):
            bc_adapter_templates = [bc_adapter_templates]
        if not bc_adapter_templates:
            raise IndexError("list index out of range")          # bc_adapter_templates[-1] in the reference
        try:
            plan = self._plan_for(qcat_config)
            subset = self._subset_for(plan, bc_adapter_templates)
        except KeyError:                                          # layouts foreign to self.layouts
            plan = self._plan_for(qcat_config, layouts=bc_adapter_templates)
            subset = list(range(len(bc_adapter_templates)))
        rec = plan.scan_windows([read_sequence or ""], subset)[0]
        result = self._record_to_dict(plan, rec)
        result["trim5p"] = 0
        result["trim3p"] = 0
        return result

    def _kits(self):
        if not self.override_kit_name:
            return self.layouts
        return self.get_adapters(self.override_kit_name)

    def _detect_records(self, plan, packed, kits):
        if plan.tables.mode == 2:
            return plan.detect(*packed)
        if not kits:
            raise IndexError("list index out of range")          # scanner_epi2me.py:64 on an empty kit list
        win5, tail3, wlen, read_len = packed
        return plan.detect(win5, tail3, wlen, read_len, self._subset_for(plan, kits))

    def _middle_found(self, kit_name, bodies, qcat_config):
        """scan_middle's answer (scanner_base.py:479-519) for many read bodies (read[W:-W], str or bytes) of one detected
        kit: True where the body or its reverse complement scans with barcode_score >= 50.

        The reference scans one read at a time and the reverse complement only when the forward scan fails; only the
        boolean is used, so here both windows of every body go to the device together, sorted by length (window
        buffers of one call are bounded to ~256 MB)."""
        detected = self.get_adapters(kit_name)
        if not detected:
            raise IndexError("list index out of range")      # bc_adapter_templates[-1] in the reference's scan()
        try:
            plan = self._plan_for(qcat_config)
            subset = self._subset_for(plan, detected)
        except KeyError:
            plan = self._plan_for(qcat_config, layouts=detected)
            subset = list(range(len(detected)))
        order = sorted(range(len(bodies)), key=lambda j: len(bodies[j]))
        found_all = [False] * len(bodies)
        start = 0
        while start < len(order):
            stop, longest = start, 16
            while stop < len(order):
                longest_next = max(longest, len(bodies[order[stop]]))
                if stop > start and 2 * (stop + 1 - start) * longest_next > (256 << 20):
                    break
                longest = longest_next
                stop += 1
            windows = []
            for j in order[start:stop]:
                body = bodies[j]
                windows.append(body)
                windows.append(body.translate(_COMP_TABLE)[::-1] if isinstance(body, bytes) else revcomp(body))
            recs = plan.scan_windows(windows, subset)
            found = ~(recs["barcode_score"] < 50.0)
            for k, j in enumerate(order[start:stop]):
                found_all[j] = bool(found[2 * k] or found[2 * k + 1])
            start = stop
        return found_all

    def _apply_middle_scan(self, read_sequences, results, qcat_config):
        """--detect-middle (scanner_base.py:593-595): reads whose body still holds an adapter of the detected kit become
        empty results with exit_status 997 (trims are kept)."""
        W = qcat_config.max_align_length
        by_kit = {}
        for i, result in enumerate(results):
            if result["adapter"]:
                by_kit.setdefault(result["adapter"].kit, []).append(i)
        for kit_name, indices in by_kit.items():
            found = self._middle_found(kit_name, [(read_sequences[i] or "")[W:-W] for i in indices], qcat_config)
            for i, hit in zip(indices, found):
                if hit:
                    trims = results[i]["trim5p"], results[i]["trim3p"]
                    results[i] = empty_return_dict()
                    results[i]["exit_status"] = 997
                    results[i]["trim5p"], results[i]["trim3p"] = trims

    def scan_middle(self, sequence, kit_name, qcat_config):
        detected = self.get_adapters(kit_name)
        W = qcat_config.max_align_length
        body = sequence[W:-W]
        for window in (body, revcomp(body)):
            middle = self.scan(window, None, detected, [], qcat_config=qcat_config)
            if middle and not middle["barcode_score"] < 50.0:
                return True
        return False

    def detect_barcode(self, read_sequence, read_qualities=None, qcat_config=None):
        """Single read (scanner_base.py:521-604)."""
        qcat_config = qcat_config or _default_config()
        plan = self._plan_for(qcat_config)
        packed = pack_windows([read_sequence], qcat_config.max_align_length)[:4]
        rec = self._detect_records(plan, packed, self._kits())[0]
        results = [self._record_to_dict(plan, rec)]
        if self.scan_middle_adapter:
            self._apply_middle_scan([read_sequence or ""], results, qcat_config)
        return results[0]

    def detect_kit(self, read_sequences, qcat_config):
        """Majority vote over the best-scoring end of every read (scanner_base.py:662-678)."""
        read_sequences = list(read_sequences)
        if not read_sequences:
            return None, []
        names = [layout.kit for layout in self.layouts]
        if len(set(names)) == 1:
            return names[0], []                     # every vote names the same kit: nothing to compute
        plan = self._plan_for(qcat_config)
        win5, tail3, wlen, _, _ = pack_windows(read_sequences, qcat_config.max_align_length)
        vote = plan.kit_vote(win5, tail3, wlen)
        return self._kit_from_votes(vote, names), []

    @staticmethod
    def _kit_from_votes(vote, names):
        # dict insertion order + stable sort by count (get_most_abundant_kits, :657-660): among kits with the
        # highest count the one seen first wins.
        counts = {}
        order = {}
        for pos, layout_index in enumerate(vote.tolist()):
            kit = names[layout_index]
            if kit not in counts:
                counts[kit] = 0
                order[kit] = pos
            counts[kit] += 1
        if not counts:
            return None
        return sorted(counts.items(), key=lambda kv: (-kv[1], order[kv[0]]))[0][0]

    def detect_barcode_batch(self, read_sequences, read_qualities=[None], qcat_config=None):
        """Batch mode (scanner_base.py:714-733): kit vote over all reads, then per-read detection restricted to
        that kit; results for zip(read_sequences, read_qualities) -- the reference's truncation is kept."""
        qcat_config = qcat_config or _default_config()
        read_sequences = list(read_sequences)
        n_out = min(len(read_sequences), len(read_qualities))
        plan = self._plan_for(qcat_config)
        packed_all = pack_windows(read_sequences, qcat_config.max_align_length)[:4]

        names = [layout.kit for layout in self.layouts]
        kit_names, kit_of_layout = plan.tables.kit_index()
        records = None
        if not read_sequences or plan.tables.mode == 2:
            kit_name = None                           # simple mode: the reference's vote never reaches scan()
        elif len(set(names)) == 1:
            kit_name = names[0]
        elif kit_of_layout is not None and hasattr(plan, "detect_auto"):
            # one pass: adapter stage over all layouts, the whole call votes as one batch on the device, detection
            # continues on the voted kit's layouts (qcb_detect_auto).  Every read votes, also those beyond n_out.
            records, batch_kit = plan.detect_auto(*packed_all, kit_of_layout, len(read_sequences), return_kits=True)
            kit_name = kit_names[int(batch_kit[0])]
            records = records[:n_out]
        else:
            kit_name = self._kit_from_votes(plan.kit_vote(packed_all[0], packed_all[1], packed_all[2]), names)

        results = []
        if n_out:
            if records is None:
                self.override_kit_name = kit_name
                try:
                    packed = tuple(a[:n_out] for a in packed_all)
                    records = self._detect_records(plan, packed, self._kits())
                finally:
                    self.override_kit_name = None
            results = self._records_to_dicts(plan, records)
            if self.scan_middle_adapter:
                self.override_kit_name = kit_name
                try:
                    self._apply_middle_scan(read_sequences[:n_out], results, qcat_config)
                finally:
                    self.override_kit_name = None

        if self.enable_filter_barcodes:
            barcode_count = {}
            for result in results:
                self.update_barcode_count(result, barcode_count)
            results = self.filter_barcodes(barcode_count, results)
        return results


_DEFAULT_CONFIG = None


def _default_config():
    # one shared default instance, like the reference's `qcat_config=config.qcatConfig()` default argument
    global _DEFAULT_CONFIG
    if _DEFAULT_CONFIG is None:
        _DEFAULT_CONFIG = config.qcatConfig()
    return _DEFAULT_CONFIG


class BarcodeScanner(GpuScannerMixin):
    """Base class with the reference's constructor and bookkeeping helpers (scanner_base.py:410-733)."""

    def __init__(self, min_quality, kit_name, kit_folder=None, enable_filter_barcodes=False,
                 scan_middle_adapter=False, device=None):
        available_kits = adapters.populate_adapter_layouts(kit_folder)
        self.min_quality = min_quality
        self.layouts = []
        self.override_kit_name = None
        self.enable_filter_barcodes = enable_filter_barcodes
        self.scan_middle_adapter = scan_middle_adapter
        self.device = device
        if kit_name and kit_name.lower() != "auto":
            self.layouts = [l for l in available_kits if kit_name.lower() == l.kit.lower()]
        else:
            self.layouts = [l for l in available_kits if l.auto_detect]

    @staticmethod
    def get_name():
        raise NotImplementedError("Abstract class")

    def get_adapters(self, kit_name):
        return [l for l in self.layouts if kit_name.lower() == l.kit.lower()]

    def get_adapter(self, kit_name):
        for layout in self.layouts:
            if kit_name.lower() == layout.kit.lower():
                return layout

    @staticmethod
    def update_kit_count(adapter, adapter_counts):
        key = adapter.kit if adapter else "none"
        adapter_counts[key] = adapter_counts.get(key, 0) + 1

    @staticmethod
    def get_most_abundant_kits(adapter_counts):
        if not adapter_counts:
            return None
        return sorted(adapter_counts.items(), key=operator.itemgetter(1), reverse=True)[0][0]

    @staticmethod
    def update_barcode_count(result, barcode_count):
        key = result["barcode"].id if result and result["barcode"] else "0"
        barcode_count[key] = barcode_count.get(key, 0) + 1

    @staticmethod
    def get_valid(barcode_counts, min_perc=0.20):
        top = max(list(barcode_counts.values()) + [0])
        min_count = int(top * min_perc)
        return [bc for bc, count in barcode_counts.items() if count > min_count]

    def filter_barcodes(self, barcode_count, results):
        """Drop barcodes seen in <= 5 % as many reads as the most abundant one (scanner_base.py:706-712)."""
        valid = self.get_valid(barcode_count, 0.05)
        for i, result in enumerate(results):
            if result and result["barcode"] and result["barcode"].id not in valid:
                results[i] = empty_return_dict()
        return results


class BarcodeScannerEPI2ME(BarcodeScanner):

    def __init__(self, min_quality=None, kit_folder=None, kit=None, enable_filter_barcodes=False,
                 scan_middle_adapter=False, threads=1, device=None):
        if min_quality is None:
            min_quality = 58
        super(BarcodeScannerEPI2ME, self).__init__(min_quality, kit, kit_folder=kit_folder,
                                                   enable_filter_barcodes=enable_filter_barcodes,
                                                   scan_middle_adapter=scan_middle_adapter, device=device)
        self.barcodes = None

    @staticmethod
    def get_name():
        return "epi2me"


class BarcodeScannerDual(BarcodeScanner):

    def __init__(self, min_quality=None, kit_folder=None, kit=None, enable_filter_barcodes=False,
                 scan_middle_adapter=False, threads=1, device=None):
        if min_quality is None:
            min_quality = 60
        super(BarcodeScannerDual, self).__init__(min_quality, "dual", kit_folder=kit_folder,
                                                 enable_filter_barcodes=enable_filter_barcodes,
                                                 scan_middle_adapter=scan_middle_adapter, device=device)
        self.barcodes = None

    @staticmethod
    def get_name():
        return "dual"


class BarcodeScannerSimple(BarcodeScanner):
    """`--simple` (reference scanner_simple.py): bare barcodes against the window, no kit knowledge."""

    def __init__(self, min_quality=None, kit_folder=None, kit=None, enable_filter_barcodes=False,
                 scan_middle_adapter=False, threads=1, device=None):
        if min_quality is None:
            min_quality = 60
        if threads != 1:
            logging.warning("Multi threading is not yet supported in simple mode. Falling back to using a single thread.")
        super(BarcodeScannerSimple, self).__init__(min_quality, None, kit_folder=kit_folder,
                                                   enable_filter_barcodes=enable_filter_barcodes,
                                                   scan_middle_adapter=scan_middle_adapter, device=device)
        if os.path.isfile(kit) and os.path.exists(kit):             # kit=None raises TypeError like the reference
            self.barcodes = adapters.get_barcodes_from_fastq(kit)
        else:
            self.barcodes = adapters.get_barcodes_simple(kit)

    @staticmethod
    def get_name():
        return "simple"

    def barcode_count(self):
        return len(self.barcodes) + 1


def get_adapter_by_name(kit, kit_folder=None):
    return [a for a in adapters.populate_adapter_layouts(kit_folder) if a.kit == kit]


def get_modes():
    return [cls.get_name() for cls in BarcodeScanner.__subclasses__()]


def get_kits(kit_folder=None):
    names = ["Auto"]
    for layout in adapters.populate_adapter_layouts(kit_folder):
        if layout.kit not in names:
            names.append(layout.kit)
    return names


def get_kits_info(kit_folder=None):
    names = {"Auto": "Auto detect kit"}
    for layout in adapters.populate_adapter_layouts(kit_folder):
        names.setdefault(layout.kit, layout.description)
    return names


def factory(mode="epi2me", min_quality=None, kit=None, kit_folder=None, enable_filter_barcodes=False,
            scan_middle_adapter=False, threads=1, device=None):
    """qcat.scanner.factory (reference scanner.py:78-111).  'guppy' falls back to epi2me as in the reference
    when pyguppy is missing; the out-of-scope mode 'brill' raises like any unknown mode."""
    if mode == "guppy":
        logging.warning("Demultiplexing mode {} currently not supported in your environment. "
                        "Falling back to epi2me.".format(mode))
        mode = "epi2me"
    for subclass in BarcodeScanner.__subclasses__():
        if mode == subclass.get_name():
            return subclass(min_quality=min_quality, kit_folder=kit_folder, kit=kit,
                            enable_filter_barcodes=enable_filter_barcodes,
                            scan_middle_adapter=scan_middle_adapter, threads=threads, device=device)
    raise RuntimeError("Invalid demultiplexing mode: {}".format(mode))
