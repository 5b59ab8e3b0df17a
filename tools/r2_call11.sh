#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
run() { # name, env...
  n=$1; shift
  env "$@" timeout 200 python bench.py $B > gpurun_out/r2c11_$n.json 2> gpurun_out/r2c11_$n.err
}
run base A=1
run worder WN_BWD6_WORDER=1
run nohint WN_L2HINT=0
run both WN_BWD6_WORDER=1 WN_L2HINT=0
WN_BWD6_WORDER=1 timeout 300 python -m pytest tests/test_gpu_fast.py -m gpu -q --timeout 300 -k "gradients or backward" 2>&1 | tail -3 > gpurun_out/r2c11_test.log
WN_TS=1 WN_BWD6_WORDER=1 timeout 200 python tools/ts_bwd.py > gpurun_out/r2c11_ts.log 2>&1
timeout 200 python tools/ts_fwd.py > gpurun_out/r2c11_tsfwd.log 2>&1
tail -n 2 gpurun_out/r2c11_test.log
python - <<'PY'
import json
for n in ("base", "worder", "nohint", "both"):
    try:
        d=json.loads(open("gpurun_out/r2c11_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:4]])
    except Exception as e:
        print(n, "failed", e)
PY
