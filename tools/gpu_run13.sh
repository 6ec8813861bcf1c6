#!/bin/bash
# v5 = v4 + noise stretches left out of the match search; gpu tests; default bench line; C3 line
O=gpurun_out/r2m
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 python tools/ab_kernels.py --mb 512 tools/ab/old.so tools/ab/v4.so tools/ab/v5.so > $O/ab.log 2>&1
grep -v "^corpora" $O/ab.log | cut -c1-230
timeout 300 python tools/stress_small_chunks.py > $O/stress.log 2>&1; tail -2 $O/stress.log
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
python - $O/bench_c2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("c2 value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, "svz", d.get("size_vs_zlib6"), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("extra_workloads",{}).items(): print(" extra", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","deflate_gbs","inflate_gbs","codec_gbs","ratio")} if isinstance(v,dict) else v)
PY
