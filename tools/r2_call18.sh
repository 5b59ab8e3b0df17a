#!/bin/bash
mkdir -p gpurun_out
STREAMS=8,64,512 timeout 300 python tools/time_gen.py > gpurun_out/r2c18_time_pipe.log 2>&1
tail -n 3 gpurun_out/r2c18_time_pipe.log
timeout 900 python -m pytest tests/test_gpu_generate.py tests/test_gpu_fullsize.py tests/test_gpu_benchshape.py -m gpu -q --timeout 400 -k "gener or stream" 2>&1 | tail -5 > gpurun_out/r2c18_test.log
tail -n 5 gpurun_out/r2c18_test.log
WN_TS=1 timeout 120 python tools/ts_gen.py 64 2>&1 | grep -v Warn | tail -13
