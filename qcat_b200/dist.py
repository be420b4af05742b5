"""Multi-GPU plumbing: reads are independent, so they are sharded across ranks (one process per GPU) with no
data-path collective; the only exchange is one all-gather of the per-barcode count vector at the end of a run
(the histogram the CLI prints, reference cli.py:386-405).  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used for that all-gather only."""
import numpy as np


def shard_indices(n_reads, rank, world_size):
    """Round-robin shard of read indices for `rank` (north star: reads sharded round-robin across GPUs)."""
    return np.arange(rank, n_reads, world_size, dtype=np.int64)


def unshard(per_rank_results, n_reads):
    """Inverse of shard_indices: interleave per-rank result arrays back into input order."""
    world = len(per_rank_results)
    out = np.zeros(n_reads, dtype=per_rank_results[0].dtype)
    for rank, res in enumerate(per_rank_results):
        out[rank::world] = res
    return out


def histogram_bins(records, layout_bin_base, n_bins):
    """Host equivalent of qcb_histogram_device: bin 0 = unclassified, 1 + base[layout] + barcode otherwise."""
    layout = records["layout"].astype(np.int64)
    barcode = records["barcode"].astype(np.int64)
    called = (layout >= 0) & (barcode >= 0)
    bins = np.zeros(len(records), dtype=np.int64)
    bins[called] = 1 + np.asarray(layout_bin_base, dtype=np.int64)[layout[called]] + barcode[called]
    return np.bincount(bins, minlength=n_bins).astype(np.int64)


def allgather_counts(counts):
    """One all_gather of the local count vector (torch tensor on the rank's device); returns the [world, n_bins]
    stack.  With world size 1 (or no process group) it is the vector itself."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts.unsqueeze(0)
    # one flat output buffer: the list form of all_gather costs ~20 ms per call with NCCL (output copies), this one < 1 ms
    world = dist.get_world_size()
    gathered = torch.zeros(world * counts.numel(), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(gathered, counts.contiguous())
    return gathered.view(world, counts.numel())


def barcode_histogram(tables, total_counts, layout_bin_base):
    """{barcode name or 'none': count} like the CLI's end-of-run histogram, from a gathered count vector."""
    out = {"none": int(total_counts[0])}
    for li, layout in enumerate(tables.layouts):
        g1 = int(tables.group[li * 2])
        if g1 < 0:
            continue
        first = tables.group_barcodes[g1]
        if tables.mode == 1:
            second = tables.group_barcodes[int(tables.group[li * 2 + 1])]
            for i, a in enumerate(first):
                for j, b in enumerate(second):
                    c = int(total_counts[1 + layout_bin_base[li] + i * len(second) + j])
                    if c:
                        key = "barcode{:02d}/{:02d}".format(a.id, b.id)
                        out[key] = out.get(key, 0) + c
        else:
            for i, a in enumerate(first):
                c = int(total_counts[1 + layout_bin_base[li] + i])
                if c:
                    out[a.name] = out.get(a.name, 0) + c
    return out
