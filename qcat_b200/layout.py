"""Adapter geometry, mirroring `qcat.layout.AdapterLayout` (reference layout.py:8-248).

An adapter template is a sequence over ATGCNX in which one or two runs of N mark where the barcode(s)
sit.  Only the accessors the detection path uses are provided, under the reference's names.
"""
import re
from collections import namedtuple

BarcodePosition = namedtuple("BarcodePosition", "start end length")

_INVALID = re.compile("[^ATGCNX]")
_PLACEHOLDER = re.compile("N+")


class AdapterLayout(object):

    def __init__(self, kit, sequence, barcode_set_1, barcode_set_2, description, auto_detect=False,
                 model=None, model_len=None, name=None, trim_offset=0):
        self.kit = kit
        self.auto_detect = auto_detect
        self.model = model
        self.model_len = model_len
        self.trim_offset = trim_offset
        self.name = name if name else kit
        self.description = description
        self.sequence = sequence.upper()
        if not self.sequence or _INVALID.search(self.sequence):
            raise RuntimeError("Invalid adapter sequence: {}".format(self.sequence))
        self.barcode_set_1 = barcode_set_1
        self.barcode_set_2 = barcode_set_2
        self.barcode_count = sum(1 for s in (barcode_set_1, barcode_set_2) if s)
        self.barcode_pos_1 = self._checked_position(barcode_set_1, 0)
        self.barcode_pos_2 = self._checked_position(barcode_set_2, 1)

    def _checked_position(self, barcode_set, index):
        if not barcode_set:
            return BarcodePosition(-1, -1, 0)
        pos = self.get_placeholder_pos(self.sequence, index)
        for barcode in barcode_set:
            if len(barcode.sequence) != pos.length:
                raise RuntimeError("Adapter length does not match place holder length: {}, {}"
                                   .format(len(barcode.sequence), pos.length))
        return pos

    @staticmethod
    def get_placeholder_pos(adapter_template, index=0):
        """Start, end (inclusive) and length of the index-th run of N (layout.py:72-96)."""
        runs = list(_PLACEHOLDER.finditer(adapter_template))
        if len(runs) > index:
            start, stop = runs[index].span()
            return BarcodePosition(start, stop - 1, stop - start)
        return BarcodePosition(-1, -1, 0)

    def __repr__(self):
        return repr({"Kit": self.kit, "Description": self.description})

    def _pos(self, index):
        if index == 0:
            return self.barcode_pos_1
        if index == 1:
            return self.barcode_pos_2
        raise RuntimeError("Invalid barcode index: {}. Must be 0 or 1 (for double barcoding)".format(index))

    def get_barcode_end(self, index=0):
        return self._pos(index).end

    def get_barcode_length(self, index=0):
        return self._pos(index).length

    def get_adapter_sequences(self, barcode_seq=None):
        if barcode_seq:
            p = self.barcode_pos_1
            return self.sequence[:p.start] + barcode_seq + self.sequence[p.end + 1:]
        return self.sequence

    def get_adapter_length(self):
        return len(self.sequence)

    def get_barcode_set(self, index=0):
        if index == 0:
            return self.barcode_set_1
        if index == 1:
            return self.barcode_set_2
        raise RuntimeError("Invalid barcode index: {}. Must be 0 or 1 (for double barcoding)".format(index))

    def get_upstream_context(self, n, index=0):
        p = self._pos(index)
        if p.end > -1:
            return self.sequence[max(0, p.start - n):p.start]
        return ""

    def get_downstream_context(self, n, index=0):
        p = self._pos(index)
        if p.end > -1:
            return self.sequence[p.end + 1:min(len(self.sequence), p.end + n + 1)]
        return ""

    def is_double_barcode(self):
        return self.barcode_set_2 is not None
