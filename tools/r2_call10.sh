#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 300 -k "gradients or backward or train_steps" 2>&1 | tail -8 > gpurun_out/r2c10_bwd6.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 200 python bench.py $B > gpurun_out/r2c10_bench_direct.json 2> gpurun_out/r2c10_bench_direct.err
WN_BWD6_TMA=1 timeout 200 python bench.py $B > gpurun_out/r2c10_bench_tma.json 2> gpurun_out/r2c10_bench_tma.err
WN_TS=1 timeout 200 python tools/ts_bwd.py > gpurun_out/r2c10_ts.log 2>&1
tail -n 4 gpurun_out/r2c10_bwd6.log
python - <<'PY'
import json
for n in ("direct", "tma"):
    try:
        d=json.loads(open("gpurun_out/r2c10_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:7]])
    except Exception as e:
        print(n, "failed", e)
PY
sed -n 1,14p gpurun_out/r2c10_ts.log | cut -c1-120
