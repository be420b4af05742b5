"""CPU, world_size 2 over gloo: round-robin sharding, per-rank histograms and the single all-gather of counts
reproduce the single-process result (the compute itself is replaced by stored golden records here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, records, base, n_bins, out_dir):
    from qcat_b200 import dist as qdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = qdist.shard_indices(len(records), rank, world)
    local = records[idx]                                    # this rank's shard of the result records
    counts = torch.from_numpy(qdist.histogram_bins(local, base, n_bins))
    gathered = qdist.allgather_counts(counts)
    assert gathered.shape == (world, n_bins)
    np.save(os.path.join(out_dir, "counts_%d.npy" % rank), gathered.sum(0).numpy())
    np.save(os.path.join(out_dir, "local_%d.npy" % rank), local)
    dist.destroy_process_group()


def test_two_rank_sharding_and_allgather(tmp_path, golden):
    from qcat_b200 import dist as qdist
    data, cases, _ = golden
    ci = [i for i, c in enumerate(cases) if c["name"] == "auto/single/all"][0]
    records = data["res_%d" % ci]
    tables, sc = helpers.tables_for_case(cases[ci])
    base = np.zeros(tables.n_layouts, dtype=np.int32)
    total = 0
    for i in range(tables.n_layouts):
        base[i] = total
        total += tables.group_size(i, 0)
    n_bins = total + 1
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), records, base, n_bins, str(tmp_path)), nprocs=world, join=True)
    want = qdist.histogram_bins(records, base, n_bins)
    for rank in range(world):
        got = np.load(tmp_path / ("counts_%d.npy" % rank))
        np.testing.assert_array_equal(got, want)
    restored = qdist.unshard([np.load(tmp_path / ("local_%d.npy" % r)) for r in range(world)], len(records))
    helpers.assert_records_equal(restored, records, "unshard")
    hist = qdist.barcode_histogram(tables, want, base)
    assert hist["none"] == int((records["barcode"] < 0).sum())
    assert sum(hist.values()) == len(records)
