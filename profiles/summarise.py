#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, committed
summaries under profiles/.

    python profiles/summarise.py launches gpurun_out/x_launches.csv  > profiles/x_launches_summary.txt
    python profiles/summarise.py report   gpurun_out/x.ncu-rep       > profiles/x_summary.txt

`launches`: per-kernel count / total / mean / share of the summed kernel time from a
`--metrics gpu__time_duration.sum --csv` pass (cold-cache, serialised: compare SHARES).
`report`: the handful of `--set full` metrics DESIGN.md and bench.py quote, per captured launch.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi])
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.1f} us of kernel time (serialised, cold cache)")
    print(f"{'kernel':64s} {'n':>5s} {'sum_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:64]:64s} {len(v):5d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:6.3f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')][:90]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:86s} {r[i]:>16s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            t = float(r[hdr.index("gpu__time_duration.sum")])
            ur, ut = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("gpu__time_duration.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[ur]
            tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(ut, 1e-6)
            uw = units[hdr.index("dram__bytes_write.sum")]
            wscale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[uw]
            traffic = rd * scale + wr * wscale
            print(f"  {'traffic = dram read + write (bytes per launch)':86s} {traffic:16.0f} byte")
            print(f"  {'traffic / duration':86s} {traffic / (t * tscale) / 1e9:16.1f} GB/s (under the profiler)")
        except Exception as e:  # noqa: BLE001
            print("  (traffic not derivable:", e, ")")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
