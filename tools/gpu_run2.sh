#!/bin/bash
# round-2 second GPU pass: new inflate on hardware — parity, bench, ncu capture of inflate_kernel
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
timeout 600 python bench.py --workload c3 --steps 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 exit $?" >> $O/bench_c3.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --files 100000"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 3 -c 1 -f -o $O/inflate $CMD > $O/ncu_inflate.log 2>&1
ZWZ_INFLATE_MODE=careful timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra > $O/bench_c2_careful.json 2> $O/bench_c2_careful.err
ls -la $O; tail -3 $O/pytest_gpu.log; tail -3 $O/bench_c2.err; python - <<'PY'
import json
for f in ("bench_c2","bench_c3","bench_c2_careful"):
    try:
        d=json.load(open(f"gpurun_out/r2b/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("inflate_gbs"))
    except Exception as e: print(f, "ERR", e)
PY
