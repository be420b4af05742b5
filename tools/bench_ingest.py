"""FASTQ -> records throughput of the native ingest path next to the GPU hot path.

  python tools/bench_ingest.py [n_reads] [mean_len]

Writes a synthetic FASTQ to /tmp, then times (a) indexing + window packing alone, (b) demux_file without outputs,
(c) demux_file with trimmed per-barcode FASTQ output.  File I/O goes through the page cache."""
import io
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from qcat_b200 import fastx, scanner, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    mean_len = float(sys.argv[2]) if len(sys.argv) > 2 else 8000.0
    sc = scanner.BarcodeScannerEPI2ME(kit="PBC096", device=0)
    data = synth.generate(sc.layouts, n, seed=4, mean_len=mean_len)
    rng = np.random.default_rng(1)
    path = os.path.join(tempfile.gettempdir(), "qcb_ingest_%d.fastq" % n)
    filler = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=60000))
    with open(path, "wb") as fh:
        for i in range(n):
            L = int(data["read_len"][i])
            seq = bytes(data["win5"][i, :150]) + filler[:L - 300] + bytes(data["tail3"][i, :150])
            fh.write(b"@read%d ch=%d\n" % (i, i % 512) + seq + b"\n+\n" + b"5" * L + b"\n")
    size = os.path.getsize(path)
    out = {"reads": n, "file_gb": size / 1e9, "cores": os.cpu_count()}

    t0 = time.perf_counter()
    total = 0
    for buf, recs, fastq in fastx.iter_chunks(path):
        fastx.pack_windows(buf, recs, 150)
        total += len(recs)
    dt = time.perf_counter() - t0
    assert total == n
    out["index_pack"] = {"reads_per_s": n / dt, "gb_per_s": size / 1e9 / dt}

    fastx.demux_file(path, sc)                                   # warm-up (plan, workspace)
    t0 = time.perf_counter()
    summary = fastx.demux_file(path, sc)
    dt = time.perf_counter() - t0
    out["demux_no_output"] = {"reads_per_s": n / dt, "gb_per_s": size / 1e9 / dt,
                              "classified": 1.0 - summary["barcodes"].get("none", 0) / n}

    outdir = tempfile.mkdtemp(prefix="qcb_out_")
    t0 = time.perf_counter()
    fastx.demux_file(path, sc, trim=True, out_dir=outdir)
    dt = time.perf_counter() - t0
    out["demux_trim_write"] = {"reads_per_s": n / dt, "gb_per_s": size / 1e9 / dt}
    print(json.dumps(out))
    os.remove(path)


if __name__ == "__main__":
    main()
