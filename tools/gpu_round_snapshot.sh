#!/bin/bash
# One GPU-box pass that produces everything profiles/ needs for a round: gpu tests, default bench (C2) both arms, C3/C1 lines,
# ncu launch list of the default bench command, one --set full capture per hot kernel on the SAME command (for dram traffic).
# usage: tools/gpu_round_snapshot.sh <tag>
TAG=${1:-round1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/env.txt; nproc >> $O/env.txt; lscpu | grep "Model name" >> $O/env.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python bench.py --workload c3 --files 2147483648 --steps 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --workload c1 --files 2147483648 --steps 3 > $O/bench_c1.json 2> $O/bench_c1.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
# warm-up = 3 steps; per step the kernels launch in the order match(4 size classes) encode md5 pack inflate md5 => skip the warm-up launches of each kernel
ncu --set full --clock-control none --import-source on -k regex:lz_match -s 12 -c 4 -f -o $O/lz_match $CMD > $O/ncu_lz_match.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:deflate_encode -s 3 -c 1 -f -o $O/deflate_encode $CMD > $O/ncu_deflate_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:inflate -s 3 -c 1 -f -o $O/inflate $CMD > $O/ncu_inflate.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md5_files -s 6 -c 1 -f -o $O/md5_files $CMD > $O/ncu_md5.log 2>&1
timeout 600 python tools/cli_compare.py --mb 2000 --ranks 2 --tmp /dev/shm --out $O/cli_compare_2GB.json > $O/cli_compare.log 2>&1
ls -la $O
