#!/usr/bin/env python
"""DRAM traffic per kernel of ONE bench step, from an ncu CSV of
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file X.csv \
       python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-e2e
The last launches of the run are the timed step (the e2e leg is off): walk back from the end until the step's kernel pattern
(lz_match x classes, deflate_encode, md5, pack, inflate, md5) is complete. Writes profiles/traffic.json when asked.
usage: tools/ncu_traffic.py X.csv [--workload c2 --files 370000 --out profiles/traffic.json]"""
import argparse
import csv
import json

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--workload", default="c2")
ap.add_argument("--files", type=int, default=370000)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.csv, errors="replace")) if len(r) > 10]
hdr = rows[0]
iid, iname, imet, ival, iunit = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = {}
order = []
for r in rows[1:]:
    k = int(r[iid])
    if k not in launches:
        launches[k] = {"name": r[iname].split("(")[0].split("::")[-1]}
        order.append(k)
    v = float(r[ival].replace(",", ""))
    u = r[iunit].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)
    launches[k][r[imet]] = v * mult
# the timed step = everything after the last-but-one inflate launch's following md5 ... simpler: the step ends with
# (inflate, md5) or inflate; walk back to the previous inflate and take what follows its trailing md5
names = [launches[k]["name"] for k in order]
inf = [i for i, n in enumerate(names) if n.startswith("inflate")]
assert len(inf) >= 2, "need at least two steps in the capture"
start = inf[-2] + 1
if start < len(names) and names[start].startswith("md5"):
    start += 1
step = order[start:]
kind = {"lz_match_kernel": "lz_match", "deflate_encode_kernel": "deflate_encode", "inflate_kernel": "inflate", "md5_files_kernel": "md5",
        "md5_files_staged_kernel": "md5", "pack_streams_kernel": "pack", "gather_records_kernel": "pack"}
agg = {}
for k in step:
    L = launches[k]
    n = kind.get(L["name"].split("<")[0], L["name"])
    d = agg.setdefault(n, {"dram_bytes": 0.0, "ms": 0.0, "launches": 0})
    d["dram_bytes"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
    d["ms"] += L.get("gpu__time_duration.sum", 0.0)
    d["launches"] += 1
out = {"workload": a.workload, "files": a.files, "how": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of one timed step",
       "kernels": {k: v["dram_bytes"] for k, v in agg.items()}, "detail": agg}
print(json.dumps(out, indent=1))
if a.out:
    json.dump(out, open(a.out, "w"), indent=1)
