#!/usr/bin/env python3
"""Turn one `ncu --set full` capture of the path's kernels into the numbers bench.py and profiles/ quote.

Recipe (GPU box, one GPU; the bench line printed under ncu is never a bench value):

    ncu --set full --clock-control none --import-source on -k regex:k_ -s 18 -c 6 -o gpurun_out/prof_r2 \
        python bench.py --hours 0.5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin
    # (0.5 h = 75 000 frames = one launch of every kernel per step; -s 18 skips the three warm-up steps)

then, here or there:

    ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv --print-units base > gpurun_out/prof_r2_raw.csv
    python tools/ncu_traffic.py gpurun_out/prof_r2_raw.csv 75000 B [profiles/ncu_r2_traffic.json]

Writes/updates the JSON bench.py reads for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per
frame and kernel) and prints a markdown table of the per-kernel counters for profiles/ncu_r2_summary.md.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("ms", "gpu__time_duration.sum", 1e-6), ("warp inst (M)", "smsp__inst_executed.sum", 1e-6),
        ("issue active %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1),
        ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
        ("LSU pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1),
        ("DRAM read (MB)", "dram__bytes_read.sum", 1e-6), ("DRAM write (MB)", "dram__bytes_write.sum", 1e-6),
        ("local load sectors (M)", "smsp__inst_executed_op_local_ld.sum", 1e-6),
        ("local store (M)", "smsp__inst_executed_op_local_st.sum", 1e-6)]


def short(name):
    m = re.search(r"(k_\w+)", name)
    return m.group(1) if m else name


def main():
    raw, frames, cfg = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    out_json = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "ncu_r2_traffic.json")
    rows = list(csv.reader(open(raw)))
    rows = [r for r in rows if len(r) > 10]
    head = rows[0]
    ix = {c: i for i, c in enumerate(head)}
    kcol = ix["Kernel Name"]
    data = [r for r in rows[2:] if r[kcol]]
    table, traffic = [], {}
    for r in data:
        k = short(r[kcol])
        vals = {}
        for label, col, scale in COLS:
            if col in ix and r[ix[col]] not in ("", "n/a"):
                vals[label] = float(r[ix[col]].replace(",", "")) * scale
        per_frame = (float(r[ix["dram__bytes_read.sum"]].replace(",", "")) + float(r[ix["dram__bytes_write.sum"]].replace(",", ""))) / frames
        traffic.setdefault(k, []).append(per_frame)
        vals["DRAM B/frame"] = per_frame
        table.append((k, vals))
    labels = [c[0] for c in COLS if any(c[0] in v for _, v in table)] + ["DRAM B/frame"]
    print("| kernel | " + " | ".join(labels) + " |")
    print("|---|" + "---|" * len(labels))
    for k, v in table:
        print("| %s | " % k + " | ".join(("%.3f" % v[c] if c == "ms" else "%.1f" % v[c]) if c in v else "" for c in labels) + " |")
    doc = {}
    if os.path.exists(out_json):
        doc = json.load(open(out_json))
    doc[cfg] = {k: round(sum(v) / len(v), 1) for k, v in traffic.items()}
    doc["source"] = "ncu --set full --clock-control none (tools/ncu_traffic.py recipe), dram__bytes_read.sum + dram__bytes_write.sum per frame"
    doc.setdefault("captures", {})[cfg] = {"raw_csv": os.path.basename(raw), "frames_per_launch": frames}
    json.dump(doc, open(out_json, "w"), indent=1, sort_keys=True)
    print("\nper frame, all kernels: %.0f bytes -> %s" % (sum(doc[cfg].values()), out_json))


if __name__ == "__main__":
    main()
