#!/bin/bash
# full gpu tests on the current build, CLI bring-up time with the process pinned to one device, ncu --set full of the three C2 kernels
O=gpurun_out/r2o
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python - <<'PY' > $O/trees.log 2>&1
import sys
sys.path.insert(0, ".")
from tools import corpus
specs, tot = [], 0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 500e6: break
    specs.append(s); tot += s.size
corpus.write_tree("/dev/shm/t_c1/w/src", specs)
PY
M=parallel-data-compression-and-decompression_b200/host/main
for rep in 1 2; do
  rm -rf /dev/shm/t_c1/arch /dev/shm/t_c1/out
  ( time ZWZ_TIMING=1 $M compress /dev/shm/t_c1/w/src /dev/shm/t_c1/arch ) > $O/cli_c1_compress_$rep.log 2>&1
  ( time ZWZ_TIMING=1 $M decompress /dev/shm/t_c1/arch /dev/shm/t_c1/out ) > $O/cli_c1_decompress_$rep.log 2>&1
done
diff -rq /dev/shm/t_c1/w/src /dev/shm/t_c1/out; echo "diff exit $?"
grep -h "real\|zwz timing\] \(compress\|decompress\)" $O/cli_c1_*.log | cut -c1-160
rm -rf /dev/shm/t_c1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-e2e"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 3 -c 1 -f -o $O/c2_inflate $CMD > $O/ncu_inflate.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:deflate_encode -s 3 -c 1 -f -o $O/c2_deflate_encode $CMD > $O/ncu_encode.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lz_match -s 12 -c 2 -f -o $O/c2_lz_match $CMD > $O/ncu_match.log 2>&1
ls -la $O | tail -12
