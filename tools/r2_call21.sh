#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py tests/test_gpu_ae.py -m gpu -q --timeout 400 -x 2>&1 | tail -3 > gpurun_out/r2c21_test.log
tail -n 3 gpurun_out/r2c21_test.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 200 python bench.py $B > gpurun_out/r2c21_fwd3.json 2> gpurun_out/r2c21_fwd3.err
WN_FWD2=1 timeout 200 python bench.py $B > gpurun_out/r2c21_fwd2.json 2> gpurun_out/r2c21_fwd2.err
python - <<'PY'
import json
for n in ("fwd3", "fwd2"):
    try:
        d=json.loads(open("gpurun_out/r2c21_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:4]])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 python tools/ae_profile.py auto 2>&1 | grep "^mode"
