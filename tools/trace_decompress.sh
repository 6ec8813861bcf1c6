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
$M compress /dev/shm/tr/w/src /dev/shm/tr/arch > /dev/null 2>&1
ZWZ_TRACE=1 ZWZ_TIMING=1 $M decompress /dev/shm/tr/arch /dev/shm/tr/out 2>&1 | grep -E "zwz trace|zwz timing" | head -60
rm -rf /dev/shm/tr/out
ZWZ_TIMING=1 ZWZ_WORKERS=1 $M decompress /dev/shm/tr/arch /dev/shm/tr/out 2>&1 | grep -E "zwz timing" | head
rm -rf /dev/shm/tr
