#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into the text files committed here.

    python profiles/summarize_ncu.py launches gpurun_out/<launches>.csv
    python profiles/summarize_ncu.py full gpurun_out/<report>.ncu-rep
"""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (cold-cache, serialised: compare shares)")
    for k, v in agg.items():
        print(f"{k:72s} launches={len(v):4d} total={sum(v) / 1e3:10.1f} us  avg={sum(v) / len(v) / 1e3:9.1f} us  share={sum(v) / tot:.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    idx = [(w, H.index(w)) for w in WANT if w in H]
    print(f"# {path}: ncu --set full, selected metrics per profiled launch")
    for r in rows[2:]:
        print("---")
        for w, i in idx:
            print(f"  {w:85s} {r[i][:60]:>20s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
