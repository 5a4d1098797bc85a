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


def _scaled(val, unit):
    """time -> ms, bytes -> MB"""
    if unit in ("ns", "nsecond"):
        return val / 1e6
    if unit in ("us", "usecond"):
        return val / 1e3
    if unit in ("ms", "msecond"):
        return val
    if unit == "second":
        return val * 1e3
    if unit == "byte":
        return val / 1e6
    if unit == "Kbyte":
        return val / 1e3
    if unit == "Mbyte":
        return val
    if unit == "Gbyte":
        return val * 1e3
    return val


def launches(path):
    """Per-kernel totals.  One metric (gpu__time_duration.sum) -> time table; with dram__bytes_* also collected,
    adds the DRAM megabytes read / written (totals over the launches listed)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    seen = collections.OrderedDict()
    for row in r:
        if len(row) < len(hdr):
            continue
        name = row[idx["Kernel Name"]].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        metric = row[idx["Metric Name"]]
        seen[metric] = 1
        val = _scaled(float(row[idx["Metric Value"]].replace(",", "")), row[idx["Metric Unit"]])
        a = agg.setdefault(name, {"ids": set()})
        a["ids"].add(row[idx["ID"]])
        a[metric] = a.get(metric, 0.0) + val
    T, R, W = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
    key = T if T in seen else R
    tot = sum(v.get(key, 0.0) for v in agg.values())
    cols = [m for m in (T, R, W) if m in seen]
    names = {T: "total ms", R: "DRAM read MB", W: "DRAM write MB"}
    print("| kernel | launches | " + " | ".join(names[c] for c in cols) + (" | share of time |" if T in seen else " |"))
    print("|---|---:|" + "---:|" * (len(cols) + (1 if T in seen else 0)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1].get(key, 0.0)):
        line = f"| `{k[:60]}` | {len(v['ids'])} | " + " | ".join(f"{v.get(c, 0.0):.3f}" for c in cols)
        if T in seen:
            line += f" | {100 * v.get(T, 0.0) / tot:.1f}%"
        print(line + " |")
    print(f"| **total** | {sum(len(v['ids']) for v in agg.values())} | " +
          " | ".join(f"{sum(v.get(c, 0.0) for v in agg.values()):.3f}" for c in cols) + (" | |" if T in seen else " |"))


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
