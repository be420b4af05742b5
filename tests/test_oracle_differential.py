"""CPU: differential test of the C oracle against the UNMODIFIED reference Python (over the parasail stand-in) on
mutated synthetic reads -- lower-case stretches, N / IUPAC / junk characters, hard truncation, homopolymer ends -- for
eight kit / mode combinations.  Complements the committed golden vectors (tests/golden), which the GPU box can check
without the reference; this one needs the reference and therefore runs where it is available."""
import os
import random
import sys

import pytest

from tests import helpers

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refloader  # noqa: E402
sys.path.pop(0)

pytestmark = pytest.mark.skipif(not refloader.available(), reason="reference package not available")


def _mutate(rng, read):
    r = list(read)
    k = rng.random()
    if k < 0.15 and r:                                   # lower-case stretch
        a = rng.randrange(0, len(r))
        b = min(len(r), a + rng.randrange(1, 80))
        r[a:b] = [c.lower() for c in r[a:b]]
    elif k < 0.3:                                        # N / IUPAC / junk
        for _ in range(rng.randrange(1, 12)):
            if r:
                r[rng.randrange(len(r))] = rng.choice("NRYKMSWBDHVnX-*")
    elif k < 0.4:                                        # hard truncation (often shorter than a barcode)
        r = r[:rng.randrange(0, 200)]
    elif k < 0.5:                                        # homopolymer in front of the adapter
        r = list(rng.choice("ACGT") * rng.randrange(1, 60)) + r
    return "".join(r)


CASES = [("epi2me", None), ("epi2me", "PBC096"), ("epi2me", "RAB204/RAB214"), ("epi2me", "NBD104/NBD114"), ("epi2me", "VMK001"),
         ("epi2me", "RPB004/RLB001"), ("epi2me", "DUAL"), ("dual", None)]


@pytest.mark.parametrize("mode,kit", CASES, ids=["%s-%s" % c for c in CASES])
def test_oracle_equals_reference_on_mutated_reads(mode, kit):
    refloader.load()
    from qcat import config as ref_config
    from qcat import scanner as ref_scanner
    from qcat_b200 import config, scanner, synth
    from qcat_b200.tables import Tables, pack_windows
    ref = ref_scanner.factory(mode=mode, kit=kit)
    mine = scanner.factory(mode=mode, kit=kit)
    assert [(l.kit, l.sequence) for l in mine.layouts] == [(l.kit, l.sequence) for l in ref.layouts]
    source = mine.layouts if (kit or mode == "dual") else scanner.factory(kit="PBK004/LWB001").layouts
    rng = random.Random(CASES.index((mode, kit)) * 7919 + 13)
    data = synth.generate(source, 260, seed=rng.randrange(1 << 20), mean_len=600.0, sub=rng.choice([0.02, 0.08, 0.12]),
                          dele=0.05, ins=0.04)
    reads = [_mutate(rng, r) for r in synth.windows_to_reads(data)]
    tables = Tables(mine.layouts, config.qcatConfig(), mode, mine.min_quality)
    win5, tail3, wlen, read_len, _ = pack_windows(reads, 150)
    got = helpers.oracle_detect(tables, win5, tail3, wlen, read_len)
    cfg = ref_config.qcatConfig()
    called = 0
    for read, g in zip(reads, got):
        w = ref.detect_barcode(read, None, cfg)
        if w["barcode"] is None:
            assert g["barcode"] < 0 and int(g["exit_status"]) == w["exit_status"], read[:80]
        else:
            called += 1
            layout = mine.layouts[int(g["layout"])]
            picked = tables.barcode_object(int(g["layout"]), int(g["barcode"]))
            ident = "{}/{}".format(picked[0].id, picked[1].id) if mode == "dual" else picked.id
            assert ident == w["barcode"].id and layout.sequence == w["adapter"].sequence, read[:80]
            assert float(g["barcode_score"]) == w["barcode_score"] and int(g["adapter_end"]) == w["adapter_end"], read[:80]
        assert (int(g["trim5p"]), int(g["trim3p"])) == (w["trim5p"], w["trim3p"]), read[:80]
    assert called > 60
