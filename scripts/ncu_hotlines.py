#!/usr/bin/env python
"""Top CUDA source lines of an ncu report by warp-stall samples (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[2] == "-" and r[0].isdigit():   # a CUDA source line aggregate row
        si = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed")
        try: lines.append((int(r[si] or 0), int(r[ii] or 0), cur, int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(l[0] for l in lines); toti = sum(l[1] for l in lines)
print(f"total samples {tot}, warp instructions {toti}")
for s, i, f, ln, src in sorted(lines, reverse=True)[:topn]:
    print(f"{100*s/tot:5.1f}% {100*i/toti:5.1f}%i  {f}:{ln:<4d} {src[:105]}")
