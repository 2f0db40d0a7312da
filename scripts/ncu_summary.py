#!/usr/bin/env python
"""Summarise ncu reports brought back in gpurun_out/ into profiles/ (tracked)."""
import csv, io, json, subprocess, sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "warps_active.avg.pct",
        "registers_per_thread", "smsp__inst_executed.sum", "issue_active.avg.pct", "pipe_fp64_cycles_active.avg.pct", "occupancy_limit",
        "sm__throughput.avg.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "Grid Size", "Block Size", "lts__t_sector_hit_rate",
        "l1tex__t_sector_hit_rate", "smsp__warp_issue_stalled", "smsp__average_warp", "achieved_occupancy", "sm__warps_active",
        "smsp__thread_inst_executed_per_inst_executed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts",
        "smsp__average_warps_issue_stalled", "memory_throughput", "sm__inst_executed_pipe_lsu", "shared")


def summarise(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    out = {}
    for i, c in enumerate(h):
        if any(k in c for k in KEYS) or c == "Kernel Name":
            out[c] = {"unit": units[i], "values": [r[i] for r in rows[2:]]}
    return out


if __name__ == "__main__":
    res = {}
    for arg in sys.argv[2:]:
        name, rep = arg.split("=")
        res[name] = summarise(rep)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
