#!/bin/bash
# round-2 fourth GPU pass: C2 with per-file content, CLI traces (where does the wall clock go), new host tests on the GPU
O=gpurun_out/r2d
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
# CLI timelines on a 2 GB C1-shaped tree and the full C2 tree, kept on /dev/shm
python - <<'PY' > $O/trees.log 2>&1
import sys, os, time
sys.path.insert(0, ".")
from tools import corpus
t=time.time()
specs, tot = [], 0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 2000e6: break
    specs.append(s); tot += s.size
corpus.write_tree("/dev/shm/t_c1/w/src", specs)
print("c1 tree", len(specs), tot, time.time()-t)
t=time.time()
corpus.write_tree("/dev/shm/t_c2/w/src", corpus.c2_specs(370000, corpus.BASE_SEED))
print("c2 tree", time.time()-t)
PY
M=parallel-data-compression-and-decompression_b200/host/main
for W in 3 4 8; do
  rm -rf /dev/shm/t_c1/arch /dev/shm/t_c1/out
  ( time ZWZ_WORKERS=$W ZWZ_TIMING=2 $M compress /dev/shm/t_c1/w/src /dev/shm/t_c1/arch ) > $O/cli_c1_compress_w$W.log 2>&1
  ( time ZWZ_WORKERS=$W ZWZ_TIMING=2 $M decompress /dev/shm/t_c1/arch /dev/shm/t_c1/out ) > $O/cli_c1_decompress_w$W.log 2>&1
done
rm -rf /dev/shm/t_c1/arch /dev/shm/t_c1/out
( time ZWZ_TRACE=1 ZWZ_TIMING=2 $M compress /dev/shm/t_c1/w/src /dev/shm/t_c1/arch ) > $O/cli_c1_compress_trace.log 2>&1
( time ZWZ_TRACE=1 ZWZ_TIMING=2 $M decompress /dev/shm/t_c1/arch /dev/shm/t_c1/out ) > $O/cli_c1_decompress_trace.log 2>&1
diff -rq /dev/shm/t_c1/w/src /dev/shm/t_c1/out > $O/cli_c1_diff.log 2>&1; echo "diff exit $?" >> $O/cli_c1_diff.log
for W in 4 8; do
  rm -rf /dev/shm/t_c2/arch /dev/shm/t_c2/out
  ( time ZWZ_WORKERS=$W ZWZ_TIMING=2 $M compress /dev/shm/t_c2/w/src /dev/shm/t_c2/arch ) > $O/cli_c2_compress_w$W.log 2>&1
  ( time ZWZ_WORKERS=$W ZWZ_TIMING=2 $M decompress /dev/shm/t_c2/arch /dev/shm/t_c2/out ) > $O/cli_c2_decompress_w$W.log 2>&1
done
( time ZWZ_WORKERS=4 ZWZ_IO_THREADS=8 ZWZ_TIMING=2 $M decompress /dev/shm/t_c2/arch /dev/shm/t_c2/out2 ) > $O/cli_c2_decompress_w4_io8.log 2>&1
( time oracle/_ref/main_ref decompress /dev/shm/t_c2/arch /dev/shm/t_c2/out3 ) 2>&1 | tail -12 > $O/cli_c2_ref_decompress_ours.log
rm -rf /dev/shm/t_c1 /dev/shm/t_c2
grep -h "real\|zwz timing\] [cd]" $O/cli_*.log | head -80
tail -3 $O/pytest_gpu.log; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d/bench_c2.json")); print(d["value"], d["e2e"]["value"], d["kernel_ms_per_step"], d.get("inflate_gbs"), d.get("ratio"), d.get("size_vs_zlib6")); print(json.dumps(d.get("extra_workloads"))[:1500])
PY
