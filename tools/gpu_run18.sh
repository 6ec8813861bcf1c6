#!/bin/bash
O=gpurun_out/r2s
mkdir -p $O
timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra --e2e-profile --e2e-workers 1 > $O/bench_w1.json 2> $O/bench_w1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2s/bench_w1.json") if l.startswith("{")][-1]
print("device", {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],1)); print(d["e2e_profile"])
PY
