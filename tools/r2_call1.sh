#!/bin/bash
# round 2, GPU call 1: full GPU suite (new bench-shape parity tests), the two-epilogue-group draft, A/B bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -40 > gpurun_out/r2c1_pytest.log
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
WN_BWD4=1 timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 300 -k "backward or gradients or train_steps" 2>&1 | tail -30 > gpurun_out/r2c1_bwd4.log
echo "bwd4 rc=$?" >> gpurun_out/r2c1_bwd4.log
WN_BWD5=1 timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_benchshape.py -m gpu -q --timeout 300 -k "backward or gradients or train_steps or padded" 2>&1 | tail -30 > gpurun_out/r2c1_bwd5.log
echo "bwd5 rc=$?" >> gpurun_out/r2c1_bwd5.log
WN_BWD5=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/r2c1_bench_bwd5.json 2> gpurun_out/r2c1_bench_bwd5.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/r2c1_bench_bwd3.json 2> gpurun_out/r2c1_bench_bwd3.err
WN_BWD4=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/r2c1_bench_bwd4.json 2> gpurun_out/r2c1_bench_bwd4.err
tail -5 gpurun_out/r2c1_pytest.log gpurun_out/r2c1_bwd4.log gpurun_out/r2c1_bwd5.log
python - <<'PY'
import json
for n in ("bwd3","bwd4","bwd5"):
    try:
        d=json.loads(open(f"gpurun_out/r2c1_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], [(k["kernel"], round(k["ms_per_step"],3)) for k in d["kernels"][:8]])
    except Exception as e:
        print(n, "failed", e)
PY
