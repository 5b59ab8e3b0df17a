#!/bin/bash
mkdir -p gpurun_out
WN_BWD6=1 timeout 300 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 200 -k "gradients or backward or train_steps" 2>&1 | tail -15 > gpurun_out/r2c4_bwd6.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
WN_BWD6=1 timeout 200 python bench.py $B > gpurun_out/r2c4_bench_bwd6.json 2> gpurun_out/r2c4_bench_bwd6.err
tail -n 6 gpurun_out/r2c4_bwd6.log
python - <<'PY'
import json
for n in ("bwd6",):
    try:
        d=json.loads(open(f"gpurun_out/r2c4_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:7]])
    except Exception as e:
        print(n, "failed", e)
PY
