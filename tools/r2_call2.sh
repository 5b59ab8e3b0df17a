#!/bin/bash
# round 2, GPU call 2: full suite (default kernels), two-group kernels (fwd3 / bwd6) parity + A/B bench, full bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2c2_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
WN_FWD3=1 WN_BWD6=1 timeout 400 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 200 -k "not generation" 2>&1 | tail -40 > gpurun_out/r2c2_new.log
echo "new rc=$?" >> gpurun_out/r2c2_new.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 200 python bench.py $B > gpurun_out/r2c2_bench_base.json 2> gpurun_out/r2c2_bench_base.err
WN_FWD3=1 timeout 200 python bench.py $B > gpurun_out/r2c2_bench_fwd3.json 2> gpurun_out/r2c2_bench_fwd3.err
WN_BWD6=1 timeout 200 python bench.py $B > gpurun_out/r2c2_bench_bwd6.json 2> gpurun_out/r2c2_bench_bwd6.err
WN_FWD3=1 WN_BWD6=1 timeout 200 python bench.py $B > gpurun_out/r2c2_bench_both.json 2> gpurun_out/r2c2_bench_both.err
timeout 900 python bench.py > gpurun_out/r2c2_bench_full.json 2> gpurun_out/r2c2_bench_full.err
tail -n 8 gpurun_out/r2c2_pytest.log gpurun_out/r2c2_new.log
python - <<'PY'
import json
for n in ("base","fwd3","bwd6","both","full"):
    try:
        d=json.loads(open(f"gpurun_out/r2c2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:6]])
    except Exception as e:
        print(n, "failed", e)
PY
