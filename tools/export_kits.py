"""Export the reference's kit definitions (qcat/resources/kits/*.yml) into qcat_b200/resources/kits.json.

Run in the build container only (needs /root/reference).  The entries are written in the order the
reference's own unsorted glob (adapters.py:144) returns the files here, because that order is what the
reference -- and therefore the golden vectors generated next to it -- use to break template ties.
"""
import glob
import json
import os
import sys

import yaml

SRC = os.path.join(os.environ.get("QCAT_REFERENCE_ROOT", "/root/reference"), "qcat", "resources", "kits")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "qcat_b200", "resources", "kits.json")


def main():
    entries = []
    for path in glob.glob(os.path.join(SRC, "*.yml")):
        with open(path) as handle:
            data = yaml.load(handle, Loader=yaml.FullLoader)
        data["file"] = os.path.basename(path)
        entries.append(data)
    with open(DST, "w") as handle:
        json.dump({"source": "nanoporetech/qcat 1.1.0 qcat/resources/kits/*.yml", "kits": entries}, handle,
                  indent=None, separators=(",", ":"), sort_keys=True)
    print("wrote %d kit entries to %s" % (len(entries), os.path.normpath(DST)))


if __name__ == "__main__":
    sys.exit(main())
