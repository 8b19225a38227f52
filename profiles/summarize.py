#!/usr/bin/env python
"""profiles/summarize.py <tag> -- turns gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_full.ncu-rep into the
tracked summaries profiles/<tag>_launches.csv (verbatim launch list), profiles/<tag>_launches.md (per-kernel
shares), profiles/<tag>_full_metrics.csv (the metrics the roofline uses, one row per captured launch) and
updates profiles/traffic.json (dram bytes per launch of the dominant kernel, read by bench.py)."""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

UNIT_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(tag, workload_note):
    src = OUT / f"{tag}_launches.csv"
    if not src.exists():
        return
    shutil.copy(src, PROF / f"{tag}_launches.csv")
    lines = [l for l in src.read_text().splitlines() if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    md = [f"# {tag}: ncu launch list summary", "", workload_note, "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).", "",
          "| launches | total ms | share | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} % | {v[1] / v[0] / 1e3:.1f} | `{k[:120]}` |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(md) + "\n")
    print("\n".join(md))


def full(tag, workload):
    rep = OUT / f"{tag}_full.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(PROF / f"{tag}_full_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
    traffic = {}
    p = PROF / "traffic.json"
    if p.exists():
        traffic = json.loads(p.read_text())
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        args = [a.strip() for a in name[name.find("<") + 1:name.rfind(">")].split(",")] if "<" in name else []
        fused_dot = ("spmv_pattern_march_kernel" in name and len(args) > 1 and args[1] == "1") or \
                    ("spmv_pattern_march_kernel" not in name and "spmv_" in name and args and args[-1] == "1")
        if fused_dot:
            rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT_TO_BYTES[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT_TO_BYTES[units[idx["dram__bytes_write.sum"]]]
            traffic[workload] = rd + wr
    p.write_text(json.dumps(traffic, indent=1) + "\n")
    print(open(PROF / f"{tag}_full_metrics.csv").read())


if __name__ == "__main__":
    tag = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "weak512"
    note = sys.argv[3] if len(sys.argv) > 3 else \
        "Command: `python bench.py --steps 1 --warmup 1 --no-also --no-cpu-baseline --no-e2e` (27-pt, 512^3, 2 solves of 149 iterations)."
    launches(tag, note)
    full(tag, workload)
