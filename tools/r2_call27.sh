#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py tests/test_gpu_ae.py tests/test_gpu_fullsize.py -m gpu -q --timeout 400 -x 2>&1 | tail -1
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 200 python bench.py $B > gpurun_out/r2c27.json 2> gpurun_out/r2c27.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c27.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:5]])
PY
timeout 200 python tools/ae_profile.py auto 2>&1 | grep "^mode"
