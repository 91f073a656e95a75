#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/: one row per profiled launch with the
counters the roofline argument uses, plus profiles/traffic.json (average DRAM bytes per launch per kernel),
which bench.py reads for roofline.traffic.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_step [--traffic]"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    recs = []
    for r in data:
        rec = {"kernel": r[name_i].split("(")[0].replace("void ", "")}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i].replace(",", "")
                if k.startswith("dram__bytes"):
                    rec[k] = to_bytes(v, units[i])
                elif k == "gpu__time_duration.sum":
                    rec["duration_ms"] = float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1.0)
                else:
                    try:
                        rec[k] = float(v)
                    except ValueError:
                        rec[k] = v
        rec["dram_bytes"] = rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0)
        if rec.get("duration_ms"):
            rec["dram_GBps"] = rec["dram_bytes"] / (rec["duration_ms"] * 1e-3) / 1e9
        recs.append(rec)
    with open(out + ".json", "w") as f:
        json.dump(recs, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write("| # | kernel | ms | DRAM read GB | DRAM write GB | DRAM GB/s | dram %% | issue %% | regs | grid |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for i, r in enumerate(recs):
            f.write("| %d | %s | %.4f | %.3f | %.3f | %.0f | %.1f | %.1f | %s | %s |\n" % (
                i, r["kernel"], r.get("duration_ms", 0), r.get("dram__bytes_read.sum", 0) / 1e9,
                r.get("dram__bytes_write.sum", 0) / 1e9, r.get("dram_GBps", 0),
                r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
                r.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
                r.get("launch__registers_per_thread", ""), r.get("launch__grid_size", "")))
    if "--traffic" in sys.argv:
        agg = {}
        for r in recs:
            base = r["kernel"].split("<")[0].split("::")[-1]
            agg.setdefault(base, []).append(r["dram_bytes"])
        tr = {k + "_dram_bytes_per_launch": sum(v) / len(v) for k, v in agg.items()}
        tr["source"] = os.path.basename(out) + ".json (ncu --set full, one bench step: 6 launches per kernel)"
        with open(os.path.join(os.path.dirname(out), "traffic.json"), "w") as f:
            json.dump(tr, f, indent=1)
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
