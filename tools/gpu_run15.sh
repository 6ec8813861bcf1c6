#!/bin/bash
O=gpurun_out/r2p
mkdir -p $O
timeout 600 python tools/e2e_sweep.py 6x32 6x16 6x24 4x16 4x32 3x12 8x32 6x48 > $O/e2e_sweep.log 2>&1; grep "^{" $O/e2e_sweep.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -2 $O/pytest_gpu.log
timeout 300 python tools/stress_small_chunks.py > $O/stress.log 2>&1; tail -1 $O/stress.log
