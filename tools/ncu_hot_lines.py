#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples and by executed instructions, from an `ncu --set full --import-source on`
report:  tools/ncu_hot_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
cur_file = ""
agg = {}
tot_s = tot_i = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] not in ("", "0") and r[0].isdigit():  # a source-line row (aggregated over its SASS)
        try:
            s, n = int(r[i_s]), int(r[i_n])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += s
        a[1] += n
        tot_s += s
        tot_i += n
print(f"total samples {tot_s}, total warp-instructions {tot_i}")
print("--- by stall samples")
for (f, ln), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}%  {100*n/max(tot_i,1):5.1f}%i  {f}:{ln}  {src[:100]}")
print("--- by instructions executed")
for (f, ln), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*n/max(tot_i,1):5.1f}%i  {100*s/max(tot_s,1):5.1f}%  {f}:{ln}  {src[:100]}")
