#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2c12_dbg.log
for m in 16 1 2 3 4 7 8 15; do
  WN_DBG=$m timeout 120 python tools/time_fwd_kernels.py 2>&1 | tail -1 >> gpurun_out/r2c12_dbg.log
done
cat gpurun_out/r2c12_dbg.log
