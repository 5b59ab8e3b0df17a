#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -x 2>&1 | grep -v "Warn\|warn" | tail -12 > gpurun_out/r2c15_test.log
tail -n 12 gpurun_out/r2c15_test.log
timeout 300 python tools/ae_profile.py auto > gpurun_out/r2c15_ae_prof.log 2>&1
grep -v Warn gpurun_out/r2c15_ae_prof.log | tail -32
timeout 300 python tools/ae_profile.py auto 4 2>&1 | grep "^mode" > gpurun_out/r2c15_ae_b4.log
cat gpurun_out/r2c15_ae_b4.log
