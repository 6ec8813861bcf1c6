#!/bin/bash
# multi-GPU box: CUDA bring-up per rank with all devices visible vs pinned to one
N=${1:-2}
O=gpurun_out/r2t
mkdir -p $O
python - <<'PY' > $O/trees.log 2>&1
import sys
sys.path.insert(0, ".")
from tools import corpus
specs, tot = [], 0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 300e6: break
    specs.append(s); tot += s.size
corpus.write_tree("/dev/shm/t_c1/w/src", specs)
PY
M=parallel-data-compression-and-decompression_b200/host/main
for mode in ${MODES:-pinned visible pinned visible}; do
  rm -rf /dev/shm/t_c1/arch /dev/shm/t_c1/out
  if [ $mode = visible ]; then export ZWZ_KEEP_DEVICES_VISIBLE=1; else unset ZWZ_KEEP_DEVICES_VISIBLE; fi
  ( time ZWZ_GPUS=$N ZWZ_TIMING=1 $M compress /dev/shm/t_c1/w/src /dev/shm/t_c1/arch ) > $O/c_$mode.log 2>&1
  ( time ZWZ_GPUS=$N ZWZ_TIMING=1 $M decompress /dev/shm/t_c1/arch /dev/shm/t_c1/out ) > $O/d_$mode.log 2>&1
  echo "$mode: compress $(grep real $O/c_$mode.log) init $(grep -o 'init [0-9.]* s' $O/c_$mode.log | tr '\n' ' ') | decompress $(grep real $O/d_$mode.log) init $(grep -o 'init [0-9.]* s' $O/d_$mode.log | tr '\n' ' ')"
  diff -rq /dev/shm/t_c1/w/src /dev/shm/t_c1/out > /dev/null; echo "diff exit $?"
done
rm -rf /dev/shm/t_c1
