#!/usr/bin/env python
"""Condense ncu output into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel totals and shares
  python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep            # key metrics per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        k = (r["Kernel Name"].split("(")[0][-60:], r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (gpu__time_duration.sum; cold-cache, serialised)")
    print("| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[0]} | {k[1]} | {k[2]} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.3f} |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    print("| kernel | | " + " | ".join(r[idx["Kernel Name"]].split("(")[0][-40:] for r in rows[2:]) + " |")
    for k in KEYS:
        if k in idx:
            print(f"| {k} | {units[idx[k]]} | " + " | ".join(r[idx[k]] for r in rows[2:]) + " |")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
