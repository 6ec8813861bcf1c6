#!/bin/bash
# Run on the GPU box (gpurun): plain run first, then the ncu launch list, then one --set full capture per hot kernel.
# Usage: tools/gpu_profile.sh <tag> "<bench args>" [kernel regexes...]
TAG=${1:-r1}; ARGS=${2:---files 20000}; shift; shift
KERNELS=${@:-lz_match deflate_encode inflate md5_files}
CMD="python bench.py $ARGS --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_$K $CMD > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out | grep ${TAG}_
