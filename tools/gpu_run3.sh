#!/bin/bash
# round-2 third GPU pass: inflate v2 (convergent decode loop, staged resolve), encoder stored regions, new compress pipeline
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
timeout 600 python bench.py --workload c3 --steps 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 exit $?" >> $O/bench_c3.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --files 100000"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 3 -c 1 -f -o $O/inflate $CMD > $O/ncu_inflate.log 2>&1
timeout 900 python tools/cli_compare.py --shape c2 --mb 2600 --ranks 8 --tmp /dev/shm --out $O/cli_compare_c2_full.json > $O/cli_compare_c2.log 2>&1
timeout 900 python tools/cli_compare.py --shape c1 --mb 2000 --ranks 2 --tmp /dev/shm --out $O/cli_compare_c1_2GB.json > $O/cli_compare_c1.log 2>&1
ls -la $O; tail -3 $O/pytest_gpu.log; tail -3 $O/bench_c2.err; python - <<'PY'
import json
for f in ("bench_c2","bench_c3"):
    try:
        d=json.load(open(f"gpurun_out/r2c/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("inflate_gbs"), d.get("ratio"), d.get("size_vs_zlib6"))
    except Exception as e: print(f, "ERR", e)
for f in ("cli_compare_c2_full","cli_compare_c1_2GB"):
    try:
        d=json.load(open(f"gpurun_out/r2c/{f}.json")); print(f, {k:d[k] for k in d if k.endswith("_s") or k.endswith("phases") or "identical" in k or k in ("ref_reads_ours","we_read_refs")})
    except Exception as e: print(f, "ERR", e)
PY
