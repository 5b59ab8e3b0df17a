#!/bin/bash
# bench.py at N GPUs the way the driver launches it (N from the environment); prints the headline of the JSON line
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_check_$N.json 2> gpurun_out/scale_check_$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_check_$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["ms_per_step"],3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "strong", d.get("strong_scaling") and d["strong_scaling"].get("ms_per_step"), "ae", d.get("autoencoder") and (round(d["autoencoder"]["ms_per_step"],3), round(d["autoencoder"]["tflops"],1)))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --impl reference --steps 2 --warmup 0 2>/dev/null | tail -c 300
