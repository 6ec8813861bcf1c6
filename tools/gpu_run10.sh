#!/bin/bash
# matcher A/B (straight-line compare + inheritance, first-inherit point) + e2e worker/part sweep
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python tools/ab_kernels.py --mb 512 tools/ab/old.so tools/ab/head1.so tools/ab/head2.so tools/ab/head4.so tools/ab/head16.so > $O/ab.log 2>&1
grep -v "^corpora" $O/ab.log | cut -c1-230
for cfg in "10 40" "12 48" "8 64"; do set -- $cfg
  timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra --e2e-workers $1 --e2e-parts $2 > $O/bench_w$1_p$2.json 2> $O/bench_w$1_p$2.err
  python - $O/bench_w$1_p$2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
PY
done
