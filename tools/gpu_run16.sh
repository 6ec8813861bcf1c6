#!/bin/bash
O=gpurun_out/r2q
mkdir -p $O
ZWZ_TRACE=1 timeout 300 python tools/e2e_sweep.py 1x32 > $O/trace_1x32.log 2> $O/trace_1x32.err
timeout 300 python tools/e2e_sweep.py 1x32 2x32 > $O/sweep_small.log 2>&1
grep "^{" $O/trace_1x32.log $O/sweep_small.log
python - <<'PY'
import re, collections
tot=collections.defaultdict(float); cnt=collections.Counter()
lines=[l for l in open("gpurun_out/r2q/trace_1x32.err") if "zwz trace" in l]
lines=lines[-2*32*4:]   # the 4 timed steps: 32 parts x 2 calls each
for l in lines:
    call=l.split("] ")[1].split(":")[0]
    for m in re.finditer(r" ([a-z0-9+ ]+?) ([0-9.]+) ms,", l):
        tot[(call,m.group(1).strip())]+=float(m.group(2)); cnt[(call,m.group(1).strip())]+=1
for k in sorted(tot): print(k, "ms per step %.1f"%(tot[k]/4))
PY
