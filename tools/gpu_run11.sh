#!/bin/bash
# matcher v3 (lazy straight-line compare, one guarded inheritance), Adler by dot products, inflate resolve MLP; e2e diagnostics
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python tools/ab_kernels.py --mb 512 tools/ab/old.so tools/ab/v3.so tools/ab/v3_oldadler.so tools/ab/v3_d12.so tools/ab/v3_mlp1.so tools/ab/v3_mlp2.so tools/ab/v3_mlp8.so > $O/ab.log 2>&1
grep -v "^corpora" $O/ab.log | cut -c1-230
for cfg in "6 32" "8 64" "8 16"; do set -- $cfg
  timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra --e2e-profile --e2e-workers $1 --e2e-parts $2 > $O/bench_w$1_p$2.json 2> $O/bench_w$1_p$2.err
  python - $O/bench_w$1_p$2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("e2e_profile"))
PY
done
