# A/B of the block backward kernels: block_bwd3 (default), the two-stage block_bwd2 (WN_BWD3=0) and - FIRST RUN PENDING - the
# two-group draft block_bwd4 (WN_BWD4=1; parity test first, everything under timeout).  Training bench only.
run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$*" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
    k={x["kernel"]: round(x["ms_per_step"], 3) for x in d["kernels"]}
    print(sys.argv[1], d["ms_per_step"], "block_bwd", k.get("block_bwd2"), "block_fwd", k.get("block_fwd"), "dx", k.get("gemm_nt_dx"), "dZcat", k.get("gemm_nt_dZcat"))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/ab.err").read()[-400:])
PY
}
run WN_BWD3=1
run WN_BWD3=0
if [ "$1" = "bwd4" ]; then
  WN_BWD4=1 timeout 120 python -m pytest tests/test_gpu_fast.py -m gpu -x -q -k "backward or train" 2>&1 | tail -3
  run WN_BWD4=1
fi
