#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) -> profiles/traffic.json:
per kernel the mean serialised duration, DRAM bytes per launch and share of the step. bench.py reads `traffic` from it.
usage: launch_list_summary.py launches.csv out.json frames_per_launch "how the list was taken" """
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mono_lidar_depth_b200.buildinfo import source_hash  # noqa: E402


def main():
    src, dst, fpl, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size")}
    per = defaultdict(dict)
    for r in rows[1:]:
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6, "byte": 1.0,
                 "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[r[ix["ID"]]][r[ix["Metric Name"]]] = v * scale
        m = re.search(r"([A-Za-z_0-9]+)(<[^(]*>)?\(", r[ix["Kernel Name"]])
        per[r[ix["ID"]]]["name"] = m.group(1) if m else r[ix["Kernel Name"]]
        per[r[ix["ID"]]]["grid"] = r[ix["Grid Size"]]
    # a sequence starts with a K1-only and ends with a gather-only fused launch, and its last chunk may be ragged: per kernel
    # only the launches with the kernel's largest grid (full launch groups) enter the per-launch figures
    def blocks(g):
        n = 1
        for t in re.findall(r"\d+", g):
            n *= int(t)
        return n

    biggest = defaultdict(int)
    for d in per.values():
        biggest[d["name"]] = max(biggest[d["name"]], blocks(d["grid"]))
    agg = defaultdict(lambda: {"n": 0, "us": 0.0, "bytes": 0.0})
    for d in per.values():
        if blocks(d["grid"]) != biggest[d["name"]]:
            continue
        a = agg[d["name"]]
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    out = {"source": how, "source_hash": source_hash(), "frames_per_launch": fpl, "share_of_step_ncu": {}}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        key = name.replace("_kernel", "")
        out[key] = {"launches": a["n"], "dram_bytes_per_launch": round(a["bytes"] / a["n"]), "avg_launch_us_ncu_serialised": round(a["us"] / a["n"], 2)}
        out["share_of_step_ncu"][key] = round(a["us"] / total, 4)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["share_of_step_ncu"]))


if __name__ == "__main__":
    main()
