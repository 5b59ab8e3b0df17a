#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
run() { n=$1; shift; env "$@" timeout 200 python bench.py $B > gpurun_out/r2c28_$n.json 2> gpurun_out/r2c28_$n.err; }
run base A=1
run noside WN_WGRAD_SIDE=0
run base2 A=1
run noside2 WN_WGRAD_SIDE=0
python - <<'PY'
import json
for n in ("base", "noside", "base2", "noside2"):
    try:
        d=json.loads(open("gpurun_out/r2c28_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:3]])
    except Exception as e:
        print(n, "failed", e)
PY
