#!/usr/bin/env python
"""Headline benchmark: reads/s demultiplexed, 96-barcode EPI2ME, 150-nt windows (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores

A step = one pass of the hot path (window orientation -> adapter DP -> template choice -> barcode DP -> two-end
decision) over one batch of synthetic reads per GPU.  `value` is measured with the batch resident in HBM (CUDA
events on the launching stream), `e2e` through the host-buffer C-ABI call (pinned host memory in, records out).
Reads are sharded across ranks (weak scaling, no data-path collective); the single NCCL all-gather of the
per-barcode counts happens once, at the end of the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_RESULT_FD = None


def emit_result(line):
    """Write the result line to the process's original stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


ALGORITHMIC_BYTES_PER_READ = 336     # 2 x 150 B windows + 4 B length in, one 32 B record out
CONFIG_INDEX = 2                     # BASELINE.json configs[2]: 96 barcodes, 1 -> 8 GPUs
KIT = "PBC096"                       # the reference's 96-barcode EPI2ME kit (NBD196 does not exist in qcat 1.1.0)
# Other BASELINE configs, selectable with --workload (parity-test cases, not the headline bench line):
WORKLOADS = {
    "configs[1]": ("NBD103/NBD104", "epi2me", "12-barcode NBD104 kit"),
    "configs[2]": ("PBC096", "epi2me", "96-barcode PBC096 kit"),
    "configs[2]-nbd196": ("NBD196", "epi2me", "synthetic 96-barcode EXP-NBD196 kit (NBD104 flanks + revcomp of the PBC096 "
                                             "barcodes, tools/make_nbd196.py)"),
    "configs[3]": (None, "dual", "dual barcoding, 24 x 96 pairs (DUAL kit)"),
    "configs[4]": ("PBC096", "epi2me", "96-barcode PBC096 kit, --trim (trim offsets are part of every record)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=1000000, help="reads per GPU per step")
    ap.add_argument("--unique-reads", type=int, default=262144, help="distinct synthetic reads generated (tiled up)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="reads timed on the CPU oracle (cpu_baseline)")
    ap.add_argument("--workload", default="configs[2]", choices=sorted(WORKLOADS))
    ap.add_argument("--kit", default=None)
    ap.add_argument("--mode", default=None, choices=["epi2me", "dual"])
    ap.add_argument("--force-generic", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    return ap.parse_args()


NBD196_FOLDER = os.path.join(ROOT, "qcat_b200", "resources", "nbd196")


def make_scanner_tables(kit, mode):
    from qcat_b200 import config, scanner
    from qcat_b200.tables import Tables
    cls = scanner.BarcodeScannerDual if mode == "dual" else scanner.BarcodeScannerEPI2ME
    # NBD196 is not a qcat kit: it is loaded from its own kit folder, like any custom kit (adapters.py:138-162)
    sc = cls(kit=kit, kit_folder=NBD196_FOLDER if kit == "NBD196" else None)
    return sc, Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality)


def synth_batch(sc, n, unique, seed):
    from qcat_b200 import scanner, synth
    foreign = scanner.BarcodeScannerEPI2ME(kit="RBK001").layouts
    u = min(n, unique)
    d = synth.generate(sc.layouts, u, seed=seed, foreign_layouts=foreign)
    reps = (n + u - 1) // u
    out = {}
    for k in ("win5", "tail3", "wlen", "read_len"):
        out[k] = np.ascontiguousarray(np.concatenate([d[k]] * reps, axis=0)[:n])
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]                  # upper half = samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_rate(tables, batch, sample, threads):
    """Reads/s of the CPU oracle (C port of the reference path, OpenMP over reads) on the first `sample` reads."""
    from tests import helpers
    n = min(sample, len(batch["wlen"]))
    sub = {k: batch[k][:n] for k in batch}
    helpers.oracle_detect(tables, sub["win5"][:256], sub["tail3"][:256], sub["wlen"][:256], sub["read_len"][:256], threads=threads)
    t0 = time.perf_counter()
    res = helpers.oracle_detect(tables, sub["win5"], sub["tail3"], sub["wlen"], sub["read_len"], threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, n, res


def reference_python_rate(args, batch, n_reads=1500):
    """Reads/s of the UNMODIFIED reference Python (qcat.scanner.factory(...).detect_barcode_batch in CLI batches of 4000,
    one process -- the reference has no other mode) over the parasail stand-in of oracle/refshim, on the first n_reads
    reads of the batch.  None when the reference package did not travel to this box.  Never raises."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refloader
        sys.path.pop(0)
        if not refloader.available():
            return None
        refloader.load()
        from qcat import config as ref_config
        from qcat import scanner as ref_scanner
        from qcat_b200 import synth
        kw = {"kit_folder": NBD196_FOLDER} if args.kit == "NBD196" else {}
        sc = ref_scanner.factory(mode=args.mode, kit=args.kit, **kw)
        n = min(n_reads, len(batch["wlen"]))
        reads = synth.windows_to_reads({k: batch[k][:n] for k in batch})
        cfg = ref_config.qcatConfig()
        sc.detect_barcode_batch(reads[:50], [None] * 50, cfg)
        t0 = time.perf_counter()
        sc.detect_barcode_batch(reads, [None] * len(reads), cfg)
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "reads/s", "cores": 1,
                "sample": "%d reads through the unmodified reference Python (detect_barcode_batch), its parasail calls served by "
                          "the scalar C stand-in of oracle/refshim" % n}
    except Exception as exc:                                   # noqa: BLE001 -- an extra number must not break the bench
        return {"unavailable": "%s: %s" % (type(exc).__name__, exc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc, tables = make_scanner_tables(args.kit, args.mode)
    threads = os.cpu_count() or 1
    batch = synth_batch(sc, args.cpu_sample, args.unique_reads, seed=20261017 + CONFIG_INDEX)
    from tests import helpers
    for _ in range(max(args.warmup, 1)):
        helpers.oracle_detect(tables, batch["win5"][:512], batch["tail3"][:512], batch["wlen"][:512], batch["read_len"][:512],
                              threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        helpers.oracle_detect(tables, batch["win5"], batch["tail3"], batch["wlen"], batch["read_len"], threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * args.cpu_sample / dt
    sample = "%d synthetic %s reads per step (bounded sample of the 10M-read config), %d steps" % (args.cpu_sample, args.kit, args.steps)
    line = {"impl": "reference", "metric": "reads/s demuxed (96-barcode EPI2ME, 150bp windows)", "value": value, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, args.cpu_sample),
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "reference = pure Python over parasail (not installable offline); this is the C oracle "
                                     "port of that path, scalar int32 affine DP, OpenMP over reads",
                             "reference_python": reference_python_rate(args, batch)},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_result(line)
    return 0


def workload_config(args, reads_per_step):
    return {"workload": "BASELINE %s: %s, %s mode, 150 nt windows, synthetic reads mean 8 kb "
                        "(8%% sub / 6%% del / 5%% ins on adapters, 10%% unbarcoded)" % (args.workload, WORKLOADS[args.workload][2], args.mode),
            "kit": args.kit, "mode": args.mode, "reads_per_gpu_per_step": reads_per_step,
            "window": 150, "sharding": "reads sharded across ranks, one all-gather of per-barcode counts at the end",
            "l2": "inputs larger than L2 (%.0f MB per step per GPU)" % (reads_per_step * 332 / 1e6)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from qcat_b200 import _ffi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    sc, tables = make_scanner_tables(args.kit, args.mode)
    plan = engine.DevicePlan(tables, device=local_rank)
    plan.set_force_generic(args.force_generic)
    n = args.reads_per_step
    batch = synth_batch(sc, n, args.unique_reads, seed=20261017 + CONFIG_INDEX + 1000 * rank)
    stride = batch["win5"].shape[1]

    d_win5 = torch.from_numpy(batch["win5"]).to(dev)
    d_tail3 = torch.from_numpy(batch["tail3"]).to(dev)
    d_wlen = torch.from_numpy(batch["wlen"]).to(dev)
    d_rlen = torch.from_numpy(batch["read_len"]).to(dev)
    d_out = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
    base, n_bins = plan.histogram_layout()
    d_counts = torch.zeros(n_bins, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        plan.detect_device(d_win5.data_ptr(), d_tail3.data_ptr(), stride, d_wlen.data_ptr(), d_rlen.data_ptr(), n,
                           d_out.data_ptr(), stream=stream)
        plan.histogram_device(d_out.data_ptr(), n, base, d_counts.data_ptr(), n_bins, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:                                          # warm the communicator outside the timed region
        dist.all_gather([torch.zeros_like(d_counts) for _ in range(world)], d_counts)
    barrier()
    d_counts.zero_()
    launches0 = plan.info()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    if world > 1:
        gathered = [torch.zeros_like(d_counts) for _ in range(world)]
        dist.all_gather(gathered, d_counts)            # the path's only collective: per-barcode counts, once
        total_counts = torch.stack(gathered).sum(0)
    else:
        total_counts = d_counts
    e1.record()
    barrier()
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = plan.info()["kernel_launches"] - launches0
    value = world * n * args.steps / (elapsed_ms * 1e-3)
    assert int(total_counts.sum().item()) == world * n * args.steps, "histogram does not account for every read"

    # ---- per-stage profile (outside the timed region): dominant kernel and its roofline -------------
    plan.set_profiling(True)
    plan.stage_times(reset=True)
    prof_steps = 2
    for _ in range(prof_steps):
        step()
    torch.cuda.synchronize()
    stages = plan.stage_times(reset=True)
    plan.set_profiling(False)
    total_stage_ms = sum(v[0] for v in stages.values()) or 1.0
    dom = max(stages, key=lambda k: stages[k][0])
    dom_ms = stages[dom][0] / prof_steps
    dom_launches = max(1, stages[dom][1] // prof_steps)              # the step runs in chunks of <= 262144 reads
    kernel_names = {"orient": "k_orient_codes", "adapter": "k_adapter_fast", "select": "k_select", "barcode": "k_barcode_fast",
                    "decide": "k_finalize", "context": "k_context"}
    traffic = None                                                    # DRAM bytes per read of that kernel, from the ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            per_read = json.load(fh).get(kernel_names.get(dom, dom), {}).get("dram_bytes_per_read")
            if per_read is not None and not args.force_generic:
                traffic = per_read * n / dom_launches
    except (OSError, ValueError):
        pass
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved_gbs = ALGORITHMIC_BYTES_PER_READ * n / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_names.get(dom, dom) if not args.force_generic else dom + " (generic)",
                "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic,
                "launches_per_step": dom_launches, "ms_per_launch": dom_ms / dom_launches,
                "algorithmic_bytes_per_launch": ALGORITHMIC_BYTES_PER_READ * n / dom_launches,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_read": ALGORITHMIC_BYTES_PER_READ,
                "kernel_ms_per_launch_set": dom_ms, "kernel_share_of_step": stages[dom][0] / total_stage_ms,
                "stage_ms_per_step": {k: v[0] / prof_steps for k, v in stages.items()},
                "note": "integer-DP path: HBM fraction is ~0 by construction (SURVEY F4); the binding roofline is "
                        "compute_roofline below"}

    # compute roofline: DP cells per second against the measured issue rate of the packed DP cell
    from tests import helpers
    ncells_sample = min(n, 20000)
    ref_cells, full = helpers.oracle_count_cells(tables, batch["win5"][:ncells_sample], batch["tail3"][:ncells_sample],
                                                 batch["wlen"][:ncells_sample])
    cells_per_read = ref_cells / ncells_sample
    cell_peak, mb_mhz = engine.microbench_cell_rate(local_rank)
    # the micro-benchmark is a 0.3 ms kernel and may run before the clocks ramp: keep its per-clock rate and evaluate
    # the peak at the SM clock sampled during the timed region
    cells_per_clk = cell_peak / (mb_mhz * 1e6)
    load_mhz = (clocks or {}).get("sm_mhz") or mb_mhz
    cell_peak = cells_per_clk * load_mhz * 1e6
    step_ms = elapsed_ms / args.steps
    compute = {"unit": "reference-equivalent DP cells/s", "cells_per_read": cells_per_read,
               "full_window_fraction": full / (2.0 * ncells_sample),
               "achieved": cells_per_read * n / (step_ms * 1e-3),
               "peak": cell_peak, "peak_source": "qcb_microbench_cell_rate (add + VIMNMX3.U16x2 per 2 cells, same box)",
               "peak_sm_mhz": load_mhz, "peak_cells_per_clk_per_sm": cells_per_clk / plan.info()["sm_count"]}
    compute["frac"] = compute["achieved"] / cell_peak
    if args.mode == "epi2me" and tables.n_groups >= 1 and len(set(int(tables.group_off[g + 1] - tables.group_off[g]) for g in range(tables.n_groups))) == 1:
        # cells the packed kernels actually execute: adapters in full, but per barcode only the 24 core columns + the
        # join, and the shared prefix / suffix once per window
        nb = int(tables.group_off[1] - tables.group_off[0])
        tlen = int(tables.tmpl_off[1] - tables.tmpl_off[0])
        adapter_cells = 2.0 * 150 * sum(int(tables.adapter_off[i + 1] - tables.adapter_off[i]) for i in range(tables.n_layouts))
        region_rows = (cells_per_read - adapter_cells) / (nb * tlen)           # summed over both windows
        kernel_cells = adapter_cells + region_rows * (nb * 25 + (tlen - 24))
        compute["executed_cells_per_read"] = kernel_cells
        compute["executed_achieved"] = kernel_cells * n / (step_ms * 1e-3)
        compute["executed_frac"] = compute["executed_achieved"] / cell_peak
        if dom == "barcode" and not args.force_generic:
            # What binds k_barcode_fast (ncu: LSU wavefronts 93 % of peak): every cell pair needs one 32-bit
            # substitution word gathered from shared memory by the lane's own base code, and the crossbar delivers 4 B
            # per lane per clock -- one wavefront per warp (32 windows x one barcode pair) per core column, plus one for
            # the row-info word.  Peak = 1 wavefront / clock / SM at the SM clock sampled under load.
            wavefronts_per_read = region_rows * (nb / 2.0) * 25.0 / 32.0
            peak_wf = plan.info()["sm_count"] * load_mhz * 1e6
            achieved_wf = wavefronts_per_read * n / (dom_ms * 1e-3)
            compute["smem_gather_roofline"] = {"kernel": "k_barcode_fast", "unit": "shared-memory wavefronts/s",
                                               "wavefronts_per_read": wavefronts_per_read, "achieved": achieved_wf,
                                               "peak": peak_wf, "frac": achieved_wf / peak_wf,
                                               "evidence": "profiles/r01d_barcode_full.md"}

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H inside) ------------
    e2e = None
    if not args.skip_e2e:
        h = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("win5", "tail3", "wlen", "read_len")}
        h_out = torch.zeros(n * 32, dtype=torch.uint8).pin_memory()
        hv = {k: v.numpy() for k, v in h.items()}
        out_view = h_out.numpy().view(_ffi.RESULT_DTYPE)
        for _ in range(2):
            plan.detect(hv["win5"], hv["tail3"], hv["wlen"], hv["read_len"], out=out_view)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.detect(hv["win5"], hv["tail3"], hv["wlen"], hv["read_len"], out=out_view)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * args.steps / float(dt.item()), "unit": "reads/s",
               "h2d_bytes_per_step": int(n * (2 * stride + 4 + 8)), "d2h_bytes_per_step": int(n * 32),
               "api": "qcb_detect (C ABI, pinned host buffers)"}
        gpu_head = out_view[:2000].copy()
    else:
        gpu_head = d_out[:2000 * 32].cpu().numpy().view(_ffi.RESULT_DTYPE)

    # ---- parity spot check + CPU baseline (rank 0) --------------------------------------------------
    cpu = None
    parity = None
    if rank == 0:
        want = helpers.oracle_detect(tables, batch["win5"][:2000], batch["tail3"][:2000], batch["wlen"][:2000], batch["read_len"][:2000])
        helpers.assert_records_equal(gpu_head, want, "bench parity spot check")
        parity = "2000 records bit-identical to the CPU oracle"
        if world == 1 and not args.skip_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, m, _ = cpu_oracle_rate(tables, batch, args.cpu_sample, threads)
            cpu = {"value": rate, "unit": "reads/s", "cores": threads, "kind": "port",
                   "sample": "first %d reads of the step batch, C oracle port (scalar int32 affine DP), OpenMP over reads" % m,
                   "reference_python": reference_python_rate(args, batch)}

    if rank == 0:
        info = plan.info()
        line = {"metric": "reads/s demuxed (96-barcode EPI2ME, 150bp windows)", "value": value, "unit": "reads/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16x2 (packed DP) / int32 / f64 scores",
                "data": "synthetic (%d unique reads tiled to %d per GPU per step)" % (min(n, args.unique_reads), n),
                "config": workload_config(args, n), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "compute_roofline": compute, "cpu_baseline": cpu, "parity": parity,
                "kernels": {"fast_adapter": info["fast_adapter"], "fast_barcode": info["fast_barcode"]}}
        emit_result(line)
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    # NCCL_DEBUG is set in the environment), so everything but the final line is sent to stderr at file-descriptor level.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    kit, mode, _ = WORKLOADS[args.workload]
    if args.mode is None:
        args.mode = mode
    if args.kit is None and args.mode != "dual":
        args.kit = kit
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
