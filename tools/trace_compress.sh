# A/B of the two arena allocators on one box: compress a 2 GB C1-shaped tree with per-call phase traces
set -e
python - <<PY
import sys, os
sys.path.insert(0, ".")
from tools import corpus
src="/dev/shm/tr/w/src"; os.makedirs(src, exist_ok=True)
specs=[]; tot=0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 2000e6: break
    specs.append(s); tot += s.size
corpus.write_tree(src, specs)
PY
M=parallel-data-compression-and-decompression_b200/host/main
for mode in 0 1 0 1; do
  rm -rf /dev/shm/tr/arch
  echo "== ZWZ_ARENA_SYNC=$mode"
  ZWZ_ARENA_SYNC=$mode ZWZ_TIMING=1 $M compress /dev/shm/tr/w/src /dev/shm/tr/arch 2>&1 | grep -E "zwz timing" | grep -E "compress:|finished"
  rm -rf /dev/shm/tr/out
  ZWZ_ARENA_SYNC=$mode ZWZ_TIMING=1 $M decompress /dev/shm/tr/arch /dev/shm/tr/out 2>&1 | grep -E "zwz timing" | grep -E "decompress:|finished"
done
rm -rf /dev/shm/tr/arch
echo "== trace (async arenas)"
ZWZ_TRACE=1 $M compress /dev/shm/tr/w/src /dev/shm/tr/arch 2>&1 | grep -E "zwz trace" | head -24
rm -rf /dev/shm/tr
