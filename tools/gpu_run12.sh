#!/bin/bash
# v4 = matcher v3 at depth 12 + inflate MLP rows with the empty-row guard; full gpu tests; e2e with sized workers
O=gpurun_out/r2l
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 python tools/ab_kernels.py --mb 512 tools/ab/old.so tools/ab/v4.so tools/ab/v4_mlp2.so tools/ab/v4_mlp3.so tools/ab/v4_mlp6.so > $O/ab.log 2>&1
grep -v "^corpora" $O/ab.log | cut -c1-230
for cfg in "6 32" "8 64" "10 64"; do set -- $cfg
  timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra --e2e-workers $1 --e2e-parts $2 > $O/bench_w$1_p$2.json 2> $O/bench_w$1_p$2.err
  python - $O/bench_w$1_p$2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d.get("e2e_profile"))
PY
done
