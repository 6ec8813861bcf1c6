#!/bin/bash
# round-2 pass after the unset-marks fix: tests, stress of the failing shape, bench lines, traffic, CLI
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python tools/stress_small_chunks.py > $O/stress.log 2>&1; tail -3 $O/stress.log
timeout 900 python bench.py --steps 5 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
timeout 600 python bench.py --workload c3 --steps 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 exit $?" >> $O/bench_c3.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-e2e"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; }
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/traffic_c2.csv $CMD > $O/ncu_traffic.log 2>&1
python tools/ncu_traffic.py $O/traffic_c2.csv --out $O/traffic.json > $O/traffic.txt 2>&1
python - <<'PY' > $O/trees.log 2>&1
import sys, os, time
sys.path.insert(0, ".")
from tools import corpus
specs, tot = [], 0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 2000e6: break
    specs.append(s); tot += s.size
corpus.write_tree("/dev/shm/t_c1/w/src", specs)
corpus.write_tree("/dev/shm/t_c2/w/src", corpus.c2_specs(370000, corpus.BASE_SEED))
PY
M=parallel-data-compression-and-decompression_b200/host/main
for T in c1 c2; do
  for rep in 1 2; do
    rm -rf /dev/shm/t_$T/arch /dev/shm/t_$T/out
    ( time ZWZ_TIMING=1 $M compress /dev/shm/t_$T/w/src /dev/shm/t_$T/arch ) > $O/cli_${T}_compress_$rep.log 2>&1
    ( time ZWZ_TIMING=1 $M decompress /dev/shm/t_$T/arch /dev/shm/t_$T/out ) > $O/cli_${T}_decompress_$rep.log 2>&1
  done
  diff -rq /dev/shm/t_$T/w/src /dev/shm/t_$T/out > $O/cli_${T}_diff.log 2>&1; echo "diff exit $?" >> $O/cli_${T}_diff.log
done
rm -rf /dev/shm/t_c1 /dev/shm/t_c2
grep -h "real" $O/cli_*.log | tr '\n' ' '; echo; cat $O/cli_*_diff.log
tail -3 $O/pytest_gpu.log; cat $O/traffic.txt | head -40; python - <<'PY'
import json
for f in ("bench_c2","bench_c3"):
    try:
        d=json.load(open(f"gpurun_out/r2h/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("inflate_gbs"), d.get("deflate_gbs"), d.get("ratio"), d.get("size_vs_zlib6"))
    except Exception as e: print(f, "ERR", e)
PY
