#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_overlap.py 2>&1 | grep -v Warn | tail -6 > gpurun_out/r2c19_overlap.log
cat gpurun_out/r2c19_overlap.log
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 2000 --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 $B > gpurun_out/r2c19_bench2.json 2> gpurun_out/r2c19_bench2.err
tail -c 1500 gpurun_out/r2c19_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c19_bench2.json").read().strip().splitlines()[-1])
    print("N=2", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "gen", d.get("generation") and round(d["generation"]["us_per_step"],2), "ae", d.get("autoencoder") and (round(d["autoencoder"]["ms_per_step"],3), round(d["autoencoder"]["tflops"],1)))
except Exception as e:
    print("failed", e)
PY
