#!/bin/bash
O=gpurun_out/r2r
mkdir -p $O
for mb in 64 128 512; do ZWZ_SWEEP_SUBBATCH_MB=$mb timeout 300 python tools/e2e_sweep.py 6x32 6x20 > $O/sweep_sb$mb.log 2>&1; echo "subbatch $mb"; grep "^{" $O/sweep_sb$mb.log; done
