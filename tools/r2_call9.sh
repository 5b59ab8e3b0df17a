#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ae_profile.py auto > gpurun_out/r2c9_ae_prof.log 2>&1
cat gpurun_out/r2c9_ae_prof.log | grep -v Warn | tail -30
