#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ae.py -m gpu -q --timeout 300 -x 2>&1 | grep -v Warn | tail -30 > gpurun_out/r2c13_ae_test.log
tail -n 30 gpurun_out/r2c13_ae_test.log
timeout 300 python tools/ae_profile.py auto > gpurun_out/r2c13_ae_prof.log 2>&1
grep -v Warn gpurun_out/r2c13_ae_prof.log | tail -40
