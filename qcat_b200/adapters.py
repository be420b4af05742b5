"""Barcode / kit model, mirroring `qcat.adapters` (reference adapters.py:15, 55-162).

Kit definitions come from (a) a folder of qcat-format YAML files (`kit_folder`, same schema as
qcat/resources/kits/*.yml) or (b) the bundled `resources/kits.json`, a consolidated export of the
reference's kit data produced by tools/export_kits.py.  The JSON keeps the layouts in the order the
reference's unsorted `glob` returned them in the build container (adapters.py:144): layout order decides
ties between equally scoring adapter templates (scanner_base.py:354), so it is part of the parity contract.
"""
import glob
import json
import logging
import os
from collections import namedtuple

from qcat_b200.layout import AdapterLayout

Barcode = namedtuple("Barcode", "name id sequence fwd_strand")

KIT_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources", "kits.json")


def read_barcode(data):
    if not data:
        return None
    return Barcode(data["name"], data["id"], data.get("sequence", None), data.get("fwd_strand", None))


def read_barcode_set(data):
    if not data:
        return None
    return [read_barcode(entry) for entry in data]


def layout_from_dict(data):
    """One kit description (YAML document or kits.json entry) -> AdapterLayout, None if inactive
    (reference adapters.py:75-105)."""
    if not data.get("active", True):
        return None
    model = model_len = None
    if data.get("model"):
        model = data["model"].get("file", None)
        model_len = data["model"].get("length", None)
    return AdapterLayout(kit=data.get("kit", ""),
                         sequence=data.get("sequence", ""),
                         barcode_set_1=read_barcode_set(data.get("barcode_set_1", None)),
                         barcode_set_2=read_barcode_set(data.get("barcode_set_2", None)),
                         description=data.get("description", ""),
                         auto_detect=data.get("auto_detect", False),
                         trim_offset=data.get("trim_offset", 0),
                         model=model, model_len=model_len)


def read_adapter_layout(filename):
    import yaml
    with open(filename, "r") as stream:
        return layout_from_dict(yaml.load(stream, Loader=yaml.FullLoader))


def _bundled_layouts():
    with open(KIT_JSON, "r") as stream:
        entries = json.load(stream)["kits"]
    return [layout for layout in (layout_from_dict(e) for e in entries) if layout]


def populate_adapter_layouts(folder=None):
    """All active layouts of a kit folder / YAML file, or the bundled set (reference adapters.py:138-162)."""
    if not folder:
        return _bundled_layouts()
    if not os.path.exists(folder):
        logging.warning("{} not found. Using default adapter sequences.".format(folder))
        return _bundled_layouts()
    filenames = glob.glob(os.path.join(folder, "*.yml")) if os.path.isdir(folder) else [folder]
    return [layout for layout in (read_adapter_layout(f) for f in filenames) if layout]


def _simple_fasta(handle):
    """(title, sequence) pairs with Bio's SimpleFastaParser semantics (what the reference parses barcode files with)."""
    title, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield title, "".join(chunks).replace(" ", "").replace("\r", "")
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.rstrip())
    if title is not None:
        yield title, "".join(chunks).replace(" ", "").replace("\r", "")


def get_barcodes_from_fastq(reads_fa):
    """Barcodes of a FASTA file, ids 1.. in file order (reference adapters.py:108-118; the name says fastq)."""
    with open(reads_fa) as handle:
        barcodes = [read_barcode({"name": title, "id": i + 1, "sequence": seq}) for i, (title, seq) in enumerate(_simple_fasta(handle))]
    if not barcodes:
        logging.error("Couldn't find barcodes in {}".format(reads_fa))
    return barcodes


def get_barcodes_simple(kit="standard", filename=None):
    """barcode_set_1 of simple_<kit>.yml (reference adapters.py:121-135): `standard` = 24, `extended` = 96 + 24 barcodes."""
    if filename and os.path.isfile(filename):
        import yaml
        with open(filename, "r") as stream:
            return read_barcode_set(yaml.load(stream, Loader=yaml.FullLoader).get("barcode_set_1", []))
    wanted = "simple_{}.yml".format(kit)
    with open(KIT_JSON, "r") as stream:
        for entry in json.load(stream)["kits"]:
            if entry.get("file") == wanted:
                return read_barcode_set(entry.get("barcode_set_1", []))
    raise IOError("[Errno 2] No such file or directory: '{}'".format(wanted))
