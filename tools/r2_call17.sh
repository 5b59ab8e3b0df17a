#!/bin/bash
mkdir -p gpurun_out
echo "LPC=2" > gpurun_out/r2c17.log
LAYERS=2,6,14,30 timeout 300 python tools/time_gen_n.py >> gpurun_out/r2c17.log 2>&1
echo "LPC=1" >> gpurun_out/r2c17.log
WN_GEN_LPC=1 LAYERS=2,6,14 timeout 300 python tools/time_gen_n.py >> gpurun_out/r2c17.log 2>&1
echo "old" >> gpurun_out/r2c17.log
WN_GEN_PIPE=0 LAYERS=2,6,14,30 timeout 300 python tools/time_gen_n.py >> gpurun_out/r2c17.log 2>&1
grep -v Warn gpurun_out/r2c17.log
