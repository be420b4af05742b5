"""Build the synthetic EXP-NBD196 kit of BASELINE configs[2] from data already in this repository.

qcat 1.1.0 has no NBD196 kit.  EXP-NBD196 is the native-barcoding chemistry of NBD104 with 96 barcodes, and the
reference's own NBD01-12 are exactly the reverse complements of its PBC01-12, so the kit is put together as the two
NBD103/NBD104 layouts (flanks, geometry, trim offset) with barcodes 1-96 = revcomp(PBC096 barcodes 1-96).  The result is
a folder of two qcat-format YAML files that both the reference (`factory(kit="NBD196", kit_folder=...)`,
adapters.py:138-162) and qcat_b200 load like any custom kit.

    python tools/make_nbd196.py [out_dir]          # default: qcat_b200/resources/nbd196
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yaml

from qcat_b200 import adapters
from qcat_b200.scanner import revcomp

DEFAULT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qcat_b200", "resources", "nbd196")


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_DIR
    layouts = adapters.populate_adapter_layouts(None)
    nbd = [l for l in layouts if l.kit == "NBD103/NBD104"]
    pbc = [l for l in layouts if l.kit == "PBC096"][0]
    assert len(nbd) == 2 and len(pbc.barcode_set_1) == 96
    barcodes = [{"name": b.name, "id": b.id, "sequence": revcomp(b.sequence), "fwd_strand": False} for b in pbc.barcode_set_1]
    for own, made in zip(nbd[0].barcode_set_1, barcodes):          # NBD01-12 really are revcomp(PBC01-12)
        assert own.sequence == made["sequence"] and own.id == made["id"]
    os.makedirs(out_dir, exist_ok=True)
    for layout in nbd:
        end = "5p" if len(layout.sequence) < 42 else "3p"          # 39-nt 5' layout, 45-nt 3' layout
        doc = {"kit": "NBD196", "description": "Synthetic EXP-NBD196: NBD104 flanks with revcomp(PBC096) barcodes 1-96 (%s)" % end,
               "active": True, "auto_detect": False, "trim_offset": layout.trim_offset, "sequence": layout.sequence,
               "barcode_set_1": barcodes}
        with open(os.path.join(out_dir, "NBD196_%s.yml" % end), "w") as handle:
            yaml.safe_dump(doc, handle, sort_keys=True)
    print("wrote NBD196_5p.yml and NBD196_3p.yml to %s" % os.path.normpath(out_dir))


if __name__ == "__main__":
    sys.exit(main())
