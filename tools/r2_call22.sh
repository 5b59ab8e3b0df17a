#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 400 -x 2>&1 | tail -1
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
run() { n=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/r2c22_$n.json 2> gpurun_out/r2c22_$n.err; }
run fwd3 A=1
run fwd3pipe WN_FWD3_PIPE=1
run fwd2 WN_FWD2=1
python - <<'PY'
import json
for n in ("fwd3", "fwd3pipe", "fwd2"):
    try:
        d=json.loads(open("gpurun_out/r2c22_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:4]])
    except Exception as e:
        print(n, "failed", e)
PY
