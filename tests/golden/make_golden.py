"""Generate tests/golden/golden_v1.npz by running the UNMODIFIED reference (qcat 1.1.0 under /root/reference)
in the build container, over the parasail stand-in of oracle/refshim (parasail itself is not installable
offline; its `sg` is restated in oracle/qcat_oracle.c).

What is stored: the two 150-nt windows + length of every input read (the detection path looks at nothing
else), and for every case the reference's result per read: layout index in scanner.layouts, barcode index
inside the layout's barcode set (dual: idx1 * n2 + idx2), barcode_score, adapter_end, trim5p, trim3p,
exit_status.  Re-run with:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refloader  # noqa: E402

qcat = refloader.load()
from qcat import config as ref_config  # noqa: E402
from qcat.scanner_epi2me import BarcodeScannerEPI2ME  # noqa: E402
from qcat.scanner_dual import BarcodeScannerDual  # noqa: E402
from qcat.test import test_barcode as ref_tests  # noqa: E402
from Bio.SeqIO.QualityIO import FastqGeneralIterator  # noqa: E402

from qcat_b200.tables import pack_windows  # noqa: E402

DATA = os.path.join(refloader.REFERENCE_ROOT, "qcat", "test", "data")
FIXTURES = ["barcode_1k.fastq", "nobarcode_1k.fastq", "nbd103.fastq", "pbk004.fastq", "rab204.fastq", "rbk004.fastq"]
RESULT_DTYPE = np.dtype([("layout", "<i4"), ("barcode", "<i4"), ("barcode_score", "<f8"), ("adapter_end", "<i4"),
                         ("trim5p", "<i4"), ("trim3p", "<i4"), ("exit_status", "<i4")], align=True)


def adversarial_reads(rng):
    """Edge cases the reference handles implicitly: empty / tiny reads, N, lower case, IUPAC, junk bytes,
    homopolymers, reads that are only a barcode, and exact copies of every adapter with a barcode filled in."""
    from qcat import adapters
    acgt = "ACGT"
    reads = ["", "A", "ACGT", "N" * 10, "N" * 200, "A" * 150, "T" * 400, "acgtacgtac" * 30,
             "ACGTRYMKVBHDN" * 20, "ACGT-*xz?" * 25, "AC GT\tAC" * 40]
    for n in (1, 2, 5, 11, 23, 24, 25, 47, 59, 60, 149, 150, 151, 299, 300, 301):
        reads.append("".join(rng.choice(acgt) for _ in range(n)))
    layouts = adapters.populate_adapter_layouts()
    for layout in layouts:
        if layout.barcode_set_2 is not None:
            continue
        for barcode in (layout.barcode_set_1[0], layout.barcode_set_1[-1]):
            full = layout.get_adapter_sequences(barcode.sequence)
            insert = "".join(rng.choice(acgt) for _ in range(rng.randint(0, 400)))
            reads.append(full + insert)                               # exact adapter at the 5' end
            reads.append("".join(rng.choice(acgt) for _ in range(rng.randint(1, 30))) + full.lower() + insert)
            reads.append(full)                                        # read == adapter only
            reads.append(barcode.sequence)                            # read == barcode only
            noisy = list(full)
            for _ in range(6):
                noisy[rng.randrange(len(noisy))] = rng.choice(acgt + "N")
            reads.append("".join(noisy) + insert + ref_tests.utils.revcomp(full) if hasattr(ref_tests, "utils") else "".join(noisy) + insert)
    # conflicting ends: barcode A at 5', reverse complement of an adapter with barcode B at 3'
    from qcat.utils import revcomp
    for layout in layouts:
        if layout.barcode_set_2 is not None or len(layout.barcode_set_1) < 2:
            continue
        a = layout.get_adapter_sequences(layout.barcode_set_1[0].sequence)
        b = layout.get_adapter_sequences(layout.barcode_set_1[1].sequence)
        insert = "".join(rng.choice(acgt) for _ in range(500))
        reads.append(a + insert + revcomp(b))
        reads.append(a + insert + revcomp(a))
    return reads


def load_fixture(name):
    with open(os.path.join(DATA, name)) as handle:
        return [seq for _, seq, _ in FastqGeneralIterator(handle)]


def encode(scanner, result):
    layout = -1
    barcode = -1
    if result["adapter"] is not None:
        layout = [i for i, l in enumerate(scanner.layouts) if l is result["adapter"]][0]
    if result["barcode"] is not None:
        L = result["adapter"]
        if scanner.get_name() == "dual":
            a, b = result["barcode"].id.split("/")
            i1 = [i for i, bc in enumerate(L.barcode_set_1) if str(bc.id) == a][0]
            i2 = [i for i, bc in enumerate(L.barcode_set_2) if str(bc.id) == b][0]
            barcode = i1 * len(L.barcode_set_2) + i2
        else:
            barcode = [i for i, bc in enumerate(L.barcode_set_1) if bc is result["barcode"]][0]
    return (layout, barcode, float(result["barcode_score"]), int(result["adapter_end"]), int(result["trim5p"]),
            int(result["trim3p"]), int(result["exit_status"]))


def build_reads():
    """The golden read set (fixtures, test literals, adversarial reads) and the index range of every part."""
    rng = random.Random(20261017)
    reads = []
    ranges = {}
    for name in FIXTURES:
        seqs = load_fixture(name)
        ranges[name] = (len(reads), len(reads) + len(seqs))
        reads += seqs
    literal = [ref_tests.read, ref_tests.read_bc3_exact, ref_tests.read_bc3, ref_tests.real_double_barcode_read]
    ranges["literals"] = (len(reads), len(reads) + len(literal))
    reads += literal
    adv = adversarial_reads(rng)
    ranges["adversarial"] = (len(reads), len(reads) + len(adv))
    reads += adv
    return reads, ranges


def small_subset(ranges):
    return ([i for i in range(ranges["barcode_1k.fastq"][0], ranges["barcode_1k.fastq"][0] + 250)] +
            [i for i in range(ranges["nobarcode_1k.fastq"][0], ranges["nobarcode_1k.fastq"][0] + 250)] +
            [i for name in FIXTURES[2:] for i in range(*ranges[name])] +
            list(range(*ranges["literals"])) + list(range(*ranges["adversarial"])))


def main():
    reads, ranges = build_reads()
    print("reads:", len(reads), ranges)

    cfg = ref_config.qcatConfig()
    cases = []
    results = {}

    def add_case(name, mode, kit, batch, indices, min_quality=None):
        cls = BarcodeScannerDual if mode == "dual" else BarcodeScannerEPI2ME
        scanner = cls(min_quality=min_quality, kit=kit)
        seqs = [reads[i] for i in indices]
        if batch:
            out = scanner.detect_barcode_batch(seqs, [None] * len(seqs), cfg)
        else:
            out = [scanner.detect_barcode(s, None, cfg) for s in seqs]
        arr = np.array([encode(scanner, r) for r in out], dtype=RESULT_DTYPE)
        key = "res_%d" % len(cases)
        results[key] = arr
        results["idx_%d" % len(cases)] = np.asarray(indices, dtype=np.int32)
        cases.append({"name": name, "mode": mode, "kit": kit, "batch": bool(batch), "min_quality": min_quality,
                      "n_layouts": len(scanner.layouts), "layout_kits": [l.kit for l in scanner.layouts]})
        called = int((arr["barcode"] >= 0).sum())
        print("%-40s reads=%5d called=%5d" % (name, len(indices), called))

    everything = list(range(len(reads)))
    fixture_all = [i for name in FIXTURES for i in range(*ranges[name])]
    small = small_subset(ranges)

    add_case("auto/single/all", "epi2me", None, False, everything)
    for kit in ("PBK004/LWB001", "RBK001", "RBK004", "NBD103/NBD104", "NBD104/NBD114", "RAB204/RAB214", "PBC096",
                "RPB004/RLB001", "VMK001", "PBC001", "DUAL"):
        add_case("%s/single/small" % kit, "epi2me", kit, False, small)
    for name in FIXTURES:
        add_case("auto/batch/%s" % name, "epi2me", None, True, list(range(*ranges[name])))
    add_case("PBC096/batch/nobarcode_1k", "epi2me", "PBC096", True, list(range(*ranges["nobarcode_1k.fastq"])))
    add_case("auto/batch/adversarial", "epi2me", None, True, list(range(*ranges["adversarial"])))
    add_case("dual/single/small", "dual", None, False, small)
    add_case("dual/batch/small", "dual", None, True, small)
    add_case("auto/single/minq0", "epi2me", None, False, small, min_quality=0)
    add_case("auto/single/minq80", "epi2me", None, False, small, min_quality=80)

    win5, tail3, wlen, read_len, stride = pack_windows(reads, cfg.max_align_length)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, win5=win5, tail3=tail3, wlen=wlen, read_len=read_len,
                        cases=np.frombuffer(json.dumps({"cases": cases, "ranges": ranges}).encode(), dtype=np.uint8),
                        **results)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
