#!/bin/bash
# multi-GPU pass: usage tools/gpu_run_multi.sh N [full]   (N = 2, 4, 8; "full" adds the copy ceiling, the C2 weak line and the CLI runs)
N=${1:-2}
O=gpurun_out/r2n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
# BASELINE configs[4]: 64 GB mixed corpus cut over N GPUs (strong scaling)
timeout 900 $TR bench.py --gpus $N --workload c5 --steps 2 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 exit $?" >> $O/bench_c5.err
python - $O/bench_c5.json <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]; print("c5", d["n_gpus"], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()}, d["config"]["workload"][:80])
except Exception as e: print("c5 ERR", e)
PY
tail -3 $O/bench_c5.err
if [ "$2" = "full" ]; then
  timeout 300 $TR tools/copy_ceiling.py --mb 1024 > $O/copy_ceiling.json 2> $O/copy_ceiling.err; cut -c1-400 $O/copy_ceiling.json
  timeout 600 $TR bench.py --gpus $N --steps 3 --no-cpu-baseline --no-extra > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
  python - $O/bench_c2.json <<'PY'
import json,sys
try:
    d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]; print("c2", d["n_gpus"], "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2))
except Exception as e: print("c2 ERR", e)
PY
  # CLI on N GPUs: C1-shaped 2 GB tree, then ONE 2 GiB text file cut over the N GPUs (the reference reads both back)
  python - <<'PY' > $O/trees.log 2>&1
import sys
sys.path.insert(0, ".")
from tools import corpus
specs, tot = [], 0
for s in corpus.c1_specs(1000, corpus.BASE_SEED):
    if tot >= 2000e6: break
    specs.append(s); tot += s.size
corpus.write_tree("/dev/shm/t_c1/w/src", specs)
import os
os.makedirs("/dev/shm/t_c3/w/src", exist_ok=True)
corpus.c3_buffer(2 << 30, 596).tofile("/dev/shm/t_c3/w/src/big.log")
PY
  M=parallel-data-compression-and-decompression_b200/host/main
  R=oracle/_ref/main_ref
  for T in c1 c3; do
    rm -rf /dev/shm/t_$T/arch /dev/shm/t_$T/out /dev/shm/t_$T/refout
    ( time ZWZ_GPUS=$N ZWZ_TIMING=1 $M compress /dev/shm/t_$T/w/src /dev/shm/t_$T/arch ) > $O/cli_${T}_compress.log 2>&1
    ls -la /dev/shm/t_$T/arch >> $O/cli_${T}_compress.log
    ( time ZWZ_GPUS=$N ZWZ_TIMING=1 $M decompress /dev/shm/t_$T/arch /dev/shm/t_$T/out ) > $O/cli_${T}_decompress.log 2>&1
    diff -rq /dev/shm/t_$T/w/src /dev/shm/t_$T/out > $O/cli_${T}_diff.log 2>&1; echo "diff(ours) exit $?" >> $O/cli_${T}_diff.log
    ( time $R decompress /dev/shm/t_$T/arch /dev/shm/t_$T/refout ) > $O/cli_${T}_refreads.log 2>&1
    diff -rq /dev/shm/t_$T/w/src /dev/shm/t_$T/refout >> $O/cli_${T}_diff.log 2>&1; echo "diff(reference reads ours) exit $?" >> $O/cli_${T}_diff.log
    echo "$T: $(grep -c 'MD5 match' $O/cli_${T}_refreads.log) match, $(grep -c 'MD5 mismatch' $O/cli_${T}_refreads.log) mismatch (reference reading our archives)" >> $O/cli_${T}_diff.log
  done
  rm -rf /dev/shm/t_c1 /dev/shm/t_c3
  grep -h real $O/cli_*.log | tr '\n' ' '; echo; cat $O/cli_*_diff.log
fi
