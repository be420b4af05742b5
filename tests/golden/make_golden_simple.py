"""Generate tests/golden/golden_simple_v1.npz: results of the UNMODIFIED reference's simple scanner
(qcat/scanner_simple.py, `--simple`) on the golden read set of make_golden.py, run in the build container over the
parasail stand-in of oracle/refshim (its sg_stats is the oracle's qo_sg_stats).

Stored per case: indices into golden_v1.npz's window arrays and one record per read -- barcode index inside the simple
barcode set (`layout` is -1: the simple scanner reports no adapter), barcode_score, adapter_end (= the best barcode's
end_query), trim5p, trim3p, exit_status.  Re-run with:  python tests/golden/make_golden_simple.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as base  # noqa: E402  (loads the reference through oracle/refloader)

from qcat import config as ref_config  # noqa: E402
from qcat.scanner_simple import BarcodeScannerSimple  # noqa: E402


def encode(scanner, result):
    barcode = -1
    if result["barcode"] is not None:
        barcode = [i for i, bc in enumerate(scanner.barcodes) if bc is result["barcode"]][0]
    assert result["adapter"] is None
    return (-1, barcode, float(result["barcode_score"]), int(result["adapter_end"]), int(result["trim5p"]),
            int(result["trim3p"]), int(result["exit_status"]))


def main():
    reads, ranges = base.build_reads()
    golden = np.load(os.path.join(HERE, "golden_v1.npz"))
    assert len(golden["wlen"]) == len(reads), "golden_v1.npz was generated from a different read set"
    cfg = ref_config.qcatConfig()
    small = base.small_subset(ranges)
    nbd = list(range(*ranges["nbd103.fastq"])) + list(range(*ranges["adversarial"]))
    cases, results = [], {}

    def add_case(name, kit, batch, indices, min_quality=None):
        scanner = BarcodeScannerSimple(min_quality=min_quality, kit=kit)
        seqs = [reads[i] for i in indices]
        if batch:
            out = scanner.detect_barcode_batch(seqs, [None] * len(seqs), cfg)
        else:
            out = [scanner.detect_barcode(s, None, cfg) for s in seqs]
        arr = np.array([encode(scanner, r) for r in out], dtype=base.RESULT_DTYPE)
        results["res_%d" % len(cases)] = arr
        results["idx_%d" % len(cases)] = np.asarray(indices, dtype=np.int32)
        cases.append({"name": name, "mode": "simple", "kit": kit, "batch": bool(batch), "min_quality": min_quality,
                      "n_barcodes": len(scanner.barcodes)})
        print("%-36s reads=%5d called=%5d" % (name, len(indices), int((arr["barcode"] >= 0).sum())))

    add_case("simple/standard/single/small", "standard", False, small)
    add_case("simple/standard/batch/small", "standard", True, small)
    add_case("simple/extended/single/nbd", "extended", False, nbd)
    add_case("simple/standard/single/minq0", "standard", False, nbd, min_quality=0)
    add_case("simple/standard/single/minq90", "standard", False, small, min_quality=90)
    out = os.path.join(HERE, "golden_simple_v1.npz")
    np.savez_compressed(out, cases=np.frombuffer(json.dumps({"cases": cases}).encode(), dtype=np.uint8), **results)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
