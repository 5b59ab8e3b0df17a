#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c26_ref.json 2> gpurun_out/r2c26_ref.err
t1=$(date +%s)
echo "reference arm wall: $((t1-t0)) s"; tail -c 700 gpurun_out/r2c26_ref.json
timeout 900 python bench.py > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err
t2=$(date +%s)
echo; echo "default bench wall: $((t2-t1)) s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c26_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["traffic"])
print(d["generation"]["us_per_step"], d["generation"]["kernel"], d["autoencoder"]["ms_per_step"], d["autoencoder"]["tflops"], d["cpu_baseline"]["value"])
PY
