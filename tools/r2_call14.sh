#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ae.py -m gpu -q --timeout 300 -x 2>&1 | grep -v "Warn\|warn" | tail -8 > gpurun_out/r2c14_ae_test.log
tail -n 8 gpurun_out/r2c14_ae_test.log
timeout 300 python tools/ae_profile.py auto > gpurun_out/r2c14_ae_prof.log 2>&1
grep -v Warn gpurun_out/r2c14_ae_prof.log | tail -40
timeout 300 python tools/ae_profile.py auto 4 2>&1 | grep "^mode" > gpurun_out/r2c14_ae_b4.log
cat gpurun_out/r2c14_ae_b4.log
