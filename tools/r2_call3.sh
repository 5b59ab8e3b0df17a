#!/bin/bash
# round 2, GPU call 3: block_bwd6 after the barrier fix, S = 512 head, fwd3 label
mkdir -p gpurun_out
WN_BWD6=1 timeout 500 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 200 -k "not generation" 2>&1 | tail -40 > gpurun_out/r2c3_bwd6.log
echo "bwd6 rc=$?" >> gpurun_out/r2c3_bwd6.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae --no-incumbent --no-cfg1 --no-dense-e2e"
WN_BWD6=1 timeout 200 python bench.py $B > gpurun_out/r2c3_bench_bwd6.json 2> gpurun_out/r2c3_bench_bwd6.err
WN_BWD6=1 WN_FWD3=1 timeout 200 python bench.py $B > gpurun_out/r2c3_bench_both.json 2> gpurun_out/r2c3_bench_both.err
timeout 300 python -m pytest tests/test_gpu_fast.py tests/test_gpu_codec.py -m gpu -q --timeout 200 -k "skip_512 or one_hot" 2>&1 | tail -40 > gpurun_out/r2c3_s512.log
tail -n 12 gpurun_out/r2c3_bwd6.log gpurun_out/r2c3_s512.log
python - <<'PY'
import json
for n in ("bwd6","both"):
    try:
        d=json.loads(open(f"gpurun_out/r2c3_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:7]])
    except Exception as e:
        print(n, "failed", e)
PY
