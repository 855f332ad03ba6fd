#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` capture of one C2 Newton step.

    python profiles/make_traffic.py gpurun_out/<dir>/c2_step.ncu-rep <src_sha16> <mesh_nodes> > profiles/r02_traffic.json

Sums dram__bytes_read.sum + dram__bytes_write.sum per launch into the phases bench.py reports (the LAST complete step of
the capture: one k_resjac_tape launch up to the next one) and stores the hash of the CUDA sources the capture was taken
on: bench.py copies the dominant phase's figure into `roofline.traffic` only when the hash and the mesh size match."""
import csv
import json
import subprocess
import sys


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    ki, gi = H.index("Kernel Name"), H.index("Grid Size") if "Grid Size" in H else H.index("launch__grid_size")
    ri, wi, ti = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def f(x):
        return float(x.replace(",", ""))
    res = []
    for r in rows[2:]:
        res.append((r[ki], r[gi], f(r[ri]) * scale[units[ri]], f(r[wi]) * scale[units[wi]], f(r[ti]), units[ti]))
    return res


def phase_of(name, seen_reduce, tail_seen):
    if "k_resjac" in name or "k_bc" in name or "k_stage_jac" in name or "k_chain_gemm" in name:
        return "residual+jacobian_blocks"
    if "k_reduce" in name:
        return "abd_reduce_level0" if not seen_reduce else "abd_reduce_upper"
    if "k_backsub" in name:
        return "abd_backsub"
    if "k_seg_cluster" in name:
        return "abd_reduce_upper"
    if "k_tail_warp" in name or "k_final" in name:
        return None  # decided by position
    return "other"


def main():
    path, sha, nodes = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = rows_of(path)
    starts = [i for i, r in enumerate(rows) if "k_resjac" in r[0]]
    lo = starts[-1]
    step = rows[lo:]
    # k_tail_warp launches: upper segments (grid > 1) before the one-block tail, back-substitution segments after it
    tail_idx = [i for i, r in enumerate(step) if "k_tail_warp" in r[0] and r[1].strip("() ").split(",")[0].strip() == "1"]
    ti = tail_idx[0] if tail_idx else None
    per = {}
    listing = []
    seen_reduce = False
    for i, (name, grid, rd, wr, dur, du) in enumerate(step):
        ph = phase_of(name, seen_reduce, ti)
        if "k_reduce" in name:
            seen_reduce = True
        if ph is None:
            ph = "abd_tail+closing_solve" if i == ti else ("abd_reduce_upper" if ti is None or i < ti else "abd_backsub")
        per[ph] = per.get(ph, 0.0) + rd + wr
        listing.append({"kernel": name[:60], "grid": grid, "phase": ph, "dram_read": rd, "dram_write": wr, "duration": dur, "duration_unit": du})
    json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` of one C2 "
                           "Newton step (n = 16, MIRK6), summed into bench.py's phases; profiles/make_traffic.py",
               "capture": path, "src_sha16": sha, "mesh_nodes": nodes,
               "dram_bytes_per_launch": {k: int(v) for k, v in per.items()}, "launches": listing}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
