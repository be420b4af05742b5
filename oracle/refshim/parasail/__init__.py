"""Stand-in for the `parasail` package (TEST INFRASTRUCTURE ONLY, build container only).

Lets the UNMODIFIED reference Python under /root/reference import and run: it binds
`parasail.sg_striped_32`, `parasail.sg_stats_striped_32`, `parasail.can_use_sse2` at import
(scanner_base.py:20-26) and `parasail.matrix_create` + pokes into `.pointer[0].matrix[i]`
(config.py:26, 245-253).  The alignment itself is the C oracle's qo_sg() / qo_sg_stats().
"""
import ctypes
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(_HERE, "..", "..")))
import build as _oracle_build  # noqa: E402  (oracle/build.py)
sys.path.pop(0)

_lib = ctypes.CDLL(_oracle_build.build())
_lib.qo_sg.restype = None
_lib.qo_sg.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                       ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(ctypes.c_uint8),
                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                       ctypes.POINTER(ctypes.c_int32)]


class _MatrixStruct(object):
    def __init__(self, size, matrix, mapper):
        self.size = size
        self.matrix = matrix      # ctypes int32 array, index-assignable
        self.mapper = mapper


class Matrix(object):
    """parasail.matrix_create(alphabet, match, mismatch): (n+1)x(n+1) row-major ints, last row/column is
    the all-zero wildcard; 256-entry case-insensitive mapper, unknown characters -> wildcard."""

    def __init__(self, alphabet, match, mismatch):
        n = len(alphabet)
        size = n + 1
        mat = (ctypes.c_int32 * (size * size))()
        for i in range(n):
            for j in range(n):
                mat[i * size + j] = match if i == j else mismatch
        mapper = (ctypes.c_uint8 * 256)(*([n] * 256))
        for i, ch in enumerate(alphabet):
            mapper[ord(ch.upper())] = i
            mapper[ord(ch.lower())] = i
        self.pointer = [_MatrixStruct(size, mat, mapper)]
        self.size = size


def matrix_create(alphabet, match, mismatch):
    return Matrix(alphabet, match, mismatch)


def can_use_sse2():
    return True


_lib.qo_sg_stats.restype = None
_lib.qo_sg_stats.argtypes = _lib.qo_sg.argtypes + [ctypes.POINTER(ctypes.c_int32)] * 3


class Result(object):
    __slots__ = ("score", "end_query", "end_ref", "matches", "similar", "length")


def _as_bytes(s):
    return s if isinstance(s, bytes) else s.encode("latin-1", "replace")


def sg(s1, s2, open, extend, matrix):
    b1, b2 = _as_bytes(s1), _as_bytes(s2)
    m = matrix.pointer[0]
    sc, eq, er = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.qo_sg(b1, len(b1), b2, len(b2), int(open), int(extend), m.matrix, m.size, m.mapper,
               ctypes.byref(sc), ctypes.byref(eq), ctypes.byref(er))
    r = Result()
    r.score, r.end_query, r.end_ref = sc.value, eq.value, er.value
    return r


sg_striped_32 = sg


def sg_stats(s1, s2, open, extend, matrix):
    """parasail.sg_stats*: used by qcat's simple scanner (find_highest_scoring_barcode(compute_identity=True)) and
    by align_adapter_identity.  Backed by the oracle's qo_sg_stats."""
    b1, b2 = _as_bytes(s1), _as_bytes(s2)
    m = matrix.pointer[0]
    out = [ctypes.c_int32() for _ in range(6)]
    _lib.qo_sg_stats(b1, len(b1), b2, len(b2), int(open), int(extend), m.matrix, m.size, m.mapper,
                     *[ctypes.byref(v) for v in out])
    r = Result()
    r.score, r.end_query, r.end_ref, r.matches, r.similar, r.length = [v.value for v in out]
    return r


sg_stats_striped_32 = sg_stats
