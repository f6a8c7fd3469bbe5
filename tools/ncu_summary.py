#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv>      -> per-kernel launch count, total time, share
  python tools/ncu_summary.py raw <report.ncu-rep>         -> key metrics per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':100s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:100]:100s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")
    print(f"{'TOTAL':100s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    ki = h.index("Kernel Name")
    for r in data:
        print("==", r[ki][:140])
        for key in KEYS:
            for i, name in enumerate(h):
                if name == key:
                    print(f"   {key:72s} {r[i]:>18s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
