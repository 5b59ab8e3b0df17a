#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ae.py -m gpu -q --timeout 400 -x 2>&1 | tail -1
timeout 200 python tools/ae_profile.py auto 2>&1 | grep -v Warn | head -12
timeout 200 python tools/ae_profile.py auto 4 2>&1 | grep "^mode"
