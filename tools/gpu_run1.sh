#!/bin/bash
# round-2 first GPU pass: baseline numbers through the new bench (plugin-call e2e), copy ceiling, full-size CLI comparisons
O=gpurun_out/r2a
mkdir -p $O
(nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv; nproc; lscpu | grep -E "Model name|NUMA"; free -g; df -h /dev/shm) > $O/env.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 120 python tools/copy_ceiling.py > $O/copy_ceiling_1gpu.json 2> $O/copy_ceiling.err
timeout 900 python bench.py --steps 5 > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --workload c3 --files 17179869184 --steps 2 --no-cpu-baseline > $O/bench_c3_16g.json 2> $O/bench_c3_16g.err; echo "c3 exit $?" >> $O/bench_c3_16g.err
timeout 900 python tools/cli_compare.py --shape c2 --mb 2600 --ranks 8 --tmp /dev/shm --out $O/cli_compare_c2_full.json > $O/cli_compare_c2.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log; cat $O/bench_c2.err | tail -5; head -c 600 $O/bench_c2.json
