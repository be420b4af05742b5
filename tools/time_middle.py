"""--detect-middle body scans (qcb_scan on long windows): row-chunked vs one-thread-per-template adapter stage.
python tools/time_middle.py [n_windows] [length]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcat_b200 import config, engine, scanner
from qcat_b200.tables import Tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
length = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
sc = scanner.BarcodeScannerEPI2ME(kit="PBC096", device=0)
plan = engine.DevicePlan(Tables(sc.layouts, config.qcatConfig(), "epi2me", sc.min_quality), device=0)
rng = np.random.default_rng(0)
windows = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=length)) for _ in range(n)]
subset = list(range(len(sc.layouts)))
for label, force in (("row-chunked", 1), ("single-thread", 2)):
    plan.set_force_generic(force)
    plan.scan_windows(windows[:64], subset)
    t0 = time.perf_counter()
    recs = plan.scan_windows(windows, subset)
    dt = time.perf_counter() - t0
    print("%s: %d windows x %d nt in %.3f s -> %.0f windows/s (incl. host packing + copies)" % (label, n, length, dt, n / dt))
