#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_gen.py > gpurun_out/r2c16_time_pipe.log 2>&1
tail -n 5 gpurun_out/r2c16_time_pipe.log
WN_GEN_PIPE=0 timeout 300 python tools/time_gen.py > gpurun_out/r2c16_time_old.log 2>&1
tail -n 3 gpurun_out/r2c16_time_old.log
timeout 900 python -m pytest tests/test_gpu_generate.py tests/test_gpu_fullsize.py tests/test_gpu_benchshape.py -m gpu -q --timeout 400 -k "gener or stream" 2>&1 | tail -15 > gpurun_out/r2c16_test.log
tail -n 15 gpurun_out/r2c16_test.log
