#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
B="--steps 20 --warmup 3 --no-cpu-baseline --gen-steps 2000 --no-incumbent --no-cfg1 --no-dense-e2e"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N $B > gpurun_out/r2c24_bench$N.json 2> gpurun_out/r2c24_bench$N.err
WN_AR_OVERLAP=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus $N $B --no-ae > gpurun_out/r2c24_bench${N}_noov.json 2> gpurun_out/r2c24_bench${N}_noov.err
python - <<'PY'
import json, os
N = os.environ.get("N", "8")
for n in ("", "_noov"):
    try:
        d=json.loads(open("gpurun_out/r2c24_bench%s%s.json" % (N, n)).read().strip().splitlines()[-1])
        print("N=" + N + n, round(d["ms_per_step"],3), "value", "%.4g" % d["value"], "e2e", round(d["e2e"]["ms_per_step"],3), "strong", d.get("strong_scaling") and {k: d["strong_scaling"][k] for k in list(d["strong_scaling"])[:4]}, "ae", d.get("autoencoder") and (round(d["autoencoder"]["ms_per_step"],3), round(d["autoencoder"]["tflops"],1)))
    except Exception as e:
        print(n, "failed", e)
PY
tail -c 400 gpurun_out/r2c24_bench$N.err
