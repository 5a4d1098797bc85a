#!/usr/bin/env python
"""Summarise ncu CSV output into the tables committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv>         per-kernel totals of gpu__time_duration
  python tools/ncu_summary.py raw <raw_page.csv>              key metrics of a `--set full` capture
                                                              (ncu -i x.ncu-rep --page raw --csv > raw_page.csv)
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for row in r:
        if len(row) < len(hdr):
            continue
        name = row[idx["Kernel Name"]].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        val = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]]
        val = val / 1e6 if unit.startswith("n") else val / 1e3 if unit.startswith("u") else val
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:60]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for row in rows[2:]:
        print(f"\n### `{row[idx['Kernel Name']][:70]}`  grid {row[idx['Grid Size']]} block {row[idx['Block Size']]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in idx:
                print(f"| {k} | {row[idx[k]]} | {units[idx[k]]} |")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
