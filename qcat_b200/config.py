"""Scoring / window configuration, mirroring `qcat.config.qcatConfig` (reference config.py:8-287).

Same attribute names, defaults and setter behaviour as the reference, but with no dependency on parasail:
the two substitution matrices are plain `ScoreMatrix` objects.  `ScoreMatrix` keeps parasail's layout
((n+1) x (n+1) row-major ints, last row/column = all-zero wildcard, case-insensitive 256-entry mapper with
unknown bytes -> wildcard) and the `.pointer[0].matrix[i]` access path, so code that pokes values the way the
reference does (config.py:245-253) keeps working.
"""
try:
    import ConfigParser
except ImportError:  # Python 3
    import configparser as ConfigParser

import numpy as np


class _MatrixView(object):
    def __init__(self, owner):
        self.matrix = owner.values      # flat numpy int32 array, index-assignable
        self.size = owner.size
        self.mapper = owner.mapper


class ScoreMatrix(object):
    """Substitution matrix with parasail.matrix_create semantics (reference config.py:26, :245)."""

    def __init__(self, alphabet, match, mismatch):
        n = len(alphabet)
        self.alphabet = alphabet
        self.size = n + 1
        m = np.zeros((self.size, self.size), dtype=np.int32)
        m[:n, :n] = mismatch
        for i in range(n):
            m[i, i] = match
        self.values = m.reshape(-1)
        self.mapper = np.full(256, n, dtype=np.uint8)
        for i, ch in enumerate(alphabet):
            self.mapper[ord(ch.upper())] = i
            self.mapper[ord(ch.lower())] = i
        self.pointer = [_MatrixView(self)]


def matrix_arrays(matrix):
    """(size, int32[size*size], uint8[256]) from a ScoreMatrix or a parasail(-like) Matrix object."""
    if isinstance(matrix, ScoreMatrix):
        return matrix.size, np.ascontiguousarray(matrix.values, dtype=np.int32), matrix.mapper.copy()
    p = matrix.pointer[0]
    size = int(p.size)
    values = np.array([int(p.matrix[i]) for i in range(size * size)], dtype=np.int32)
    mapper = np.array([int(p.mapper[i]) for i in range(256)], dtype=np.uint8)
    return size, values, mapper


class qcatConfig(object):
    """Drop-in for qcat.config.qcatConfig (reference config.py:8)."""

    def __init__(self, config_path=None):
        self._match = 5
        self._nmatch = -1
        self._mismatch = -2
        self._gap_open = 2
        self._gap_extend = 2
        self._max_align_length = 150
        self._extracted_barcode_extension = 11
        self._barcode_context_length = 11
        self.matrix = None
        self.update_matrix()
        self._matrix_barcode = ScoreMatrix("ATGCN", 1, -1)      # config.py:26
        if config_path is not None:
            self.read(config_path)

    @property
    def matrix_barcode(self):
        return self._matrix_barcode

    # The setters normalise signs exactly like the reference (config.py:39-129).
    @property
    def match(self):
        return self._match

    @match.setter
    def match(self, value):
        self._match = abs(value)
        self.update_matrix()

    @property
    def nmatch(self):
        return self._nmatch

    @nmatch.setter
    def nmatch(self, value):
        self._nmatch = abs(value)
        self.update_matrix()

    @property
    def mismatch(self):
        return self._mismatch

    @mismatch.setter
    def mismatch(self, value):
        self._mismatch = -1 * abs(value)
        self.update_matrix()

    @property
    def gap_open(self):
        return self._gap_open

    @gap_open.setter
    def gap_open(self, value):
        self._gap_open = abs(value)

    @property
    def gap_extend(self):
        return self._gap_extend

    @gap_extend.setter
    def gap_extend(self, value):
        self._gap_extend = abs(value)

    @property
    def max_align_length(self):
        return self._max_align_length

    @max_align_length.setter
    def max_align_length(self, value):
        self._max_align_length = value

    @property
    def extracted_barcode_extension(self):
        return self._extracted_barcode_extension

    @extracted_barcode_extension.setter
    def extracted_barcode_extension(self, value):
        self._extracted_barcode_extension = value

    @property
    def barcode_context_length(self):
        return self._barcode_context_length

    @barcode_context_length.setter
    def barcode_context_length(self, value):
        self._barcode_context_length = value

    def update_matrix(self):
        """Adapter matrix over "ATGCNX" (+wildcard): N row/column = nmatch, X row/column = 0 (config.py:236-253)."""
        self.matrix = ScoreMatrix("ATGCNX", self.match, self.mismatch)
        flat = self.matrix.pointer[0].matrix
        for i in (4, 11, 18, 25, 28, 29, 30, 31, 32):
            flat[i] = self.nmatch
        for i in (5, 12, 19, 26, 33, 35, 36, 37, 38, 39, 40):
            flat[i] = 0

    _INT_KEYS = ("gap_open", "gap_extend", "match", "mismatch", "max_align_length",
                 "extracted_barcode_extension", "barcode_context_length")

    def write(self, out_config_path):
        parser = ConfigParser.RawConfigParser()
        parser.add_section("qcat")
        for key in self._INT_KEYS:
            parser.set("qcat", key, str(getattr(self, key)))
        with open(out_config_path, "w") as handle:
            parser.write(handle)

    def read(self, config_path):
        parser = ConfigParser.RawConfigParser()
        parser.read(config_path)
        for key in self._INT_KEYS:
            setattr(self, key, parser.getint("qcat", key))


def get_default_config():
    return qcatConfig()
