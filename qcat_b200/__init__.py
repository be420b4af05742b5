"""qcat_b200 -- B200-native (sm_100a) engine for qcat's EPI2ME / dual barcode-detection hot path.

Public surface (mirrors nanoporetech/qcat): `qcat_b200.scanner.factory`, `BarcodeScannerEPI2ME`,
`BarcodeScannerDual` with `detect_barcode`, `detect_barcode_batch`, `scan`; `qcat_b200.config.qcatConfig`;
`qcat_b200.dropin.install()` to accelerate an installed qcat in place.  The compute lives in
libqcat_b200.so (qcat_b200/csrc, C ABI in include/qcat_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
