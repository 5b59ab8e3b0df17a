#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ae.py -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r2c8_ae.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_ae.py 2>&1 | tail -30 > gpurun_out/r2c8_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-incumbent --no-cfg1 --no-dense-e2e > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
tail -n 25 gpurun_out/r2c8_ae.log; tail -n 6 gpurun_out/r2c8_pytest.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c8_bench.json").read().strip().splitlines()[-1])
    print("step", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:7]])
    print("AE", d["autoencoder"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2c8_bench.err").read()[-2000:])
PY
