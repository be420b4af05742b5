"""FASTQ -> records throughput of the native ingest path next to the GPU hot path.

  python tools/bench_ingest.py [n_reads] [mean_len]

Writes a synthetic FASTQ to /tmp, then times (a) reading + indexing + window packing alone (1 thread / all cores),
(b) demux_file without outputs, (c) with the TSV table, (d) with trimmed per-barcode FASTQ output, (e) kit auto.
File I/O goes through the page cache."""
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

    threads = os.cpu_count()
    for label, thr in (("index_pack_1thread", 1), ("index_pack", threads)):
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            total = 0
            with fastx.Reader(path, 64 << 20, thr) as reader:
                for chunk in reader.chunks(1):
                    fastx.pack_windows(chunk.data, chunk.recs, 150, thr)
                    total += len(chunk)
                    chunk.release()
            best = min(best, time.perf_counter() - t0)
        assert total == n
        out[label] = {"reads_per_s": n / best, "gb_per_s": size / 1e9 / best, "threads": thr}

    def timed(**kw):
        fastx.demux_file(path, sc, keep_records=False, **kw)        # warm-up (plan, workspace, page cache)
        t0 = time.perf_counter()
        summary = fastx.demux_file(path, sc, keep_records=False, **kw)
        dt = time.perf_counter() - t0
        return summary, {"reads_per_s": n / dt, "gb_per_s": size / 1e9 / dt}

    summary, out["demux_no_output"] = timed()
    out["demux_no_output"]["classified"] = 1.0 - summary["barcodes"].get("none", 0) / n
    with open(os.devnull, "wb") as sink:
        _, out["demux_tsv"] = timed(tsv=sink)
    outdir = tempfile.mkdtemp(prefix="qcb_out_")
    _, out["demux_trim_write"] = timed(trim=True, out_dir=outdir)
    auto = scanner.BarcodeScannerEPI2ME(device=0)                   # kit auto: per-batch vote over 12 layouts
    t0 = time.perf_counter()
    fastx.demux_file(path, auto, keep_records=False)
    t0 = time.perf_counter()
    fastx.demux_file(path, auto, keep_records=False)
    dt = time.perf_counter() - t0
    out["demux_auto_kit"] = {"reads_per_s": n / dt, "gb_per_s": size / 1e9 / dt}
    out["demux_auto_kit_by_chunk_mb"] = {}
    for mb in (64, 128, 256):
        t0 = time.perf_counter()
        fastx.demux_file(path, auto, keep_records=False, chunk_bytes=mb << 20)
        dt = time.perf_counter() - t0
        out["demux_auto_kit_by_chunk_mb"][str(mb)] = size / 1e9 / dt
    print(json.dumps(out))
    os.remove(path)


if __name__ == "__main__":
    main()
