#!/bin/bash
mkdir -p gpurun_out
WN_TS=1 WN_BWD6=1 timeout 200 python tools/ts_bwd.py > gpurun_out/r2c5_ts.log 2>&1
cat gpurun_out/r2c5_ts.log
