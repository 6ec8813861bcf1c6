#!/bin/bash
# end-of-round snapshot on one B200: gpu tests, both bench arms, C3 at 16 GiB, launch list + per-kernel DRAM traffic of the default bench command
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/env.txt; nproc >> $O/env.txt; lscpu | grep "Model name" >> $O/env.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "c2 exit $?" >> $O/bench_c2.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload c3 --files 17179869184 --steps 2 --no-cpu-baseline > $O/bench_c3_16g.json 2> $O/bench_c3_16g.err; echo "c3 exit $?" >> $O/bench_c3_16g.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-e2e"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; }
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_traffic_c2.csv $CMD > $O/ncu_traffic.log 2>&1
python tools/ncu_traffic.py $O/launches_traffic_c2.csv --out $O/traffic.json > $O/traffic.txt 2>&1
python - <<'PY'
import json
def last(f): return [json.loads(l) for l in open(f) if l.startswith("{")][-1]
for f in ("bench_c2","bench_c3_16g","bench_reference"):
    try:
        d=last(f"gpurun_out/final/{f}.json"); print(f, "value", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,1) for k,v in d.get("kernel_ms_per_step",{}).items()}, "svz", d.get("size_vs_zlib6"), "roof", d.get("roofline",{}).get("frac"))
        for k,v in d.get("extra_workloads",{}).items(): print("   extra", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ("value","deflate_gbs","inflate_gbs","codec_gbs","ratio")} if isinstance(v,dict) else v)
    except Exception as e: print(f, "ERR", e)
PY
head -30 $O/traffic.txt
