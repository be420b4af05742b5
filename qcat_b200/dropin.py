"""Accelerate an installed nanoporetech/qcat in place.

`install()` rebinds the detection entry points of qcat's own `BarcodeScannerEPI2ME`, `BarcodeScannerDual` and
`BarcodeScannerSimple`
(detect_barcode, detect_barcode_batch, scan, scan_middle, detect_kit) to the GPU-backed implementations of
`qcat_b200.scanner.GpuScannerMixin`.  Everything else -- the constructors (so `self.layouts` keeps the
reference's own kit loading and ordering), `qcat.scanner.factory`, `qcat.cli` -- stays the reference's code,
which is what "drops in under qcat/cli.py unchanged" means (cli.py:477-513 only calls factory(),
detect_barcode() and detect_barcode_batch()).  `uninstall()` restores the original methods.

    import qcat_b200.dropin; qcat_b200.dropin.install()
    from qcat import cli; cli.main(["-f", "reads.fastq", "-k", "PBC096", "-b", "out/"])
"""
from qcat_b200.scanner import GpuScannerMixin

_saved = {}


def _targets():
    import qcat.scanner  # noqa: F401  (import order: scanner <-> scanner_epi2me cycle, see scanner.py:13)
    from qcat.scanner_dual import BarcodeScannerDual
    from qcat.scanner_epi2me import BarcodeScannerEPI2ME
    from qcat.scanner_simple import BarcodeScannerSimple
    return [BarcodeScannerEPI2ME, BarcodeScannerDual, BarcodeScannerSimple]


def install(device=None, devices=None):
    """Patch qcat's scanner classes; `device` pins the CUDA device index for every scanner created afterwards,
    `devices` (a list of indices or "all") spreads every batch over several GPUs of this process."""
    if devices is not None:
        device = devices
    for cls in _targets():
        if cls in _saved:
            continue
        saved = {}
        for name, attr in GpuScannerMixin.__dict__.items():
            if name.startswith("__") and name.endswith("__"):
                continue
            saved[name] = cls.__dict__.get(name, _MISSING)
            setattr(cls, name, attr)
        if device is not None:
            cls.device = device
        _saved[cls] = saved
    return True


def uninstall():
    for cls, saved in list(_saved.items()):
        for name, attr in saved.items():
            if attr is _MISSING:
                if name in cls.__dict__:
                    delattr(cls, name)
            else:
                setattr(cls, name, attr)
        del _saved[cls]


class _Missing(object):
    pass


_MISSING = _Missing()
