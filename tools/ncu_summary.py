"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rNN_kernel_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_selected.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as out:
        out.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        out.write("source: %s\n\n| kernel | launches | total ms | ms / launch | share |\n|---|---|---|---|---|\n" % src)
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| %s | %d | %.3f | %.3f | %.1f %% |\n" % (name[:70], a[0], a[1], a[1] / a[0], 100 * a[1] / total))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write("ncu --set full --clock-control none --import-source on; source: %s\n" % src)
        for rec in rows[2:]:
            name = rec[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            out.write("\nkernel: %s\n\n| metric | value | unit |\n|---|---|---|\n" % name[:100])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.write("| %s | %s | %s |\n" % (k, rec[i], units[i]))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
