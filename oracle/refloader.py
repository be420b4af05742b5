"""Import the UNMODIFIED reference (build container only; /root/reference does not exist on the GPU box).

TEST INFRASTRUCTURE ONLY.  Puts the parasail/Bio stand-ins (oracle/refshim) and /root/reference on sys.path
and returns the `qcat` package.  Used by tests/golden/make_golden.py and by the tests that compare the C
oracle with the reference's own Python orchestration.
"""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "refshim")
# /root/reference exists only in the build container; baseline/_ref is the same package pip-installed
# (unmodified, `pip install --no-deps --target baseline/_ref`), git-ignored but shipped to the GPU box.
_CANDIDATES = [os.environ.get("QCAT_REFERENCE_ROOT", ""), "/root/reference",
               os.path.normpath(os.path.join(_HERE, "..", "baseline", "_ref"))]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "qcat"))), "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qcat"))


def load():
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import qcat.scanner  # noqa: F401  (scanner first: scanner <-> scanner_epi2me import cycle)
        import qcat
    return qcat
