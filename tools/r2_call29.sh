#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py tests/test_gpu_fullsize.py -m gpu -q --timeout 400 -x 2>&1 | tail -1
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
run() { n=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/r2c29_$n.json 2> gpurun_out/r2c29_$n.err; }
run new A=1
run tma WN_BWD6_TMA=1
run new2 A=1
python - <<'PY'
import json
for n in ("new", "tma", "new2"):
    try:
        d=json.loads(open("gpurun_out/r2c29_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:3]])
    except Exception as e:
        print(n, "failed", e)
PY
WN_TS=1 timeout 200 python tools/ts_bwd.py 2>&1 | sed -n 1,12p | cut -c1-118
