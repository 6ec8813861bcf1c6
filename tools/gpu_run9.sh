#!/bin/bash
# matcher A/B (walk/extend split, inheritance, good-length cut) + phase trace of the end-to-end leg
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python tools/ab_kernels.py --mb 512 tools/ab/old.so tools/ab/walk3.so tools/ab/walk4.so tools/ab/walk6.so tools/ab/walk8.so tools/ab/walk16.so tools/ab/walk6_good16.so tools/ab/walk6_good8.so > $O/ab.log 2>&1
grep -v "^corpora" $O/ab.log | cut -c1-260
ZWZ_TRACE=1 timeout 600 python bench.py --steps 2 --no-cpu-baseline --no-extra > $O/bench_trace.json 2> $O/bench_trace.err
grep -c "zwz trace" $O/bench_trace.err
# ncu of the two kernels that bound C3 (the product build = walk6)
CMD="python bench.py --workload c3 --files 536870912 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/c3_plain.log 2>&1 || { echo "plain c3 run failed"; tail -5 $O/c3_plain.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz_match -s 3 -c 1 -f -o $O/c3_lz_match $CMD > $O/ncu_c3_lz_match.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 3 -c 1 -f -o $O/c3_inflate $CMD > $O/ncu_c3_inflate.log 2>&1
ls -la $O
