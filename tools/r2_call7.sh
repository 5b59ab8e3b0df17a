#!/bin/bash
mkdir -p gpurun_out
WN_BWD6=1 WN_BWD6_DIRECT=1 timeout 300 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 200 -k "gradients or backward or train_steps" 2>&1 | tail -8 > gpurun_out/r2c7_bwd6.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
WN_BWD6=1 WN_BWD6_DIRECT=1 timeout 200 python bench.py $B > gpurun_out/r2c7_bench_bwd6.json 2> gpurun_out/r2c7_bench_bwd6.err
WN_TS=1 WN_BWD6=1 WN_BWD6_DIRECT=1 timeout 200 python tools/ts_bwd.py > gpurun_out/r2c7_ts.log 2>&1
tail -n 4 gpurun_out/r2c7_bwd6.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c7_bench_bwd6.json").read().strip().splitlines()[-1])
    print("bwd6", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:7]])
except Exception as e:
    print("failed", e)
PY
sed -n 8,14p gpurun_out/r2c7_ts.log
