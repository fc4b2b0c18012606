#!/usr/bin/env python3
"""Summarises an .ncu-rep (ncu --set full) into the JSON committed under profiles/: one record per profiled launch with
the metrics the roofline discussion uses. Usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.json "command" "note" """
import csv, io, json, subprocess, sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__inst_executed_pipe_fp64.sum"]


def main():
    rep, out, command, note = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        rec = {}
        for w in WANT:
            if w in head:
                k = head.index(w)
                rec[w] = r[k]
                if units[k] and w != "Kernel Name":
                    rec[w + " [unit]"] = units[k]
        launches.append(rec)
    json.dump({"command": command, "note": note, "launches": launches}, open(out, "w"), indent=1)
    for l in launches:
        print(l["Kernel Name"][:60], l.get("gpu__time_duration.sum"), l.get("dram__bytes_read.sum"), l.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
