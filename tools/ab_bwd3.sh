# A/B of the three-stage block backward kernel (default) against the two-stage block_bwd2 (WN_BWD3=0): training bench only.
run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$*" <<'PY'
import json, sys
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
k={x["kernel"]: round(x["ms_per_step"], 3) for x in d["kernels"]}
print(sys.argv[1], d["ms_per_step"], "block_bwd", k.get("block_bwd2"), "block_fwd", k.get("block_fwd"), "dx", k.get("gemm_nt_dx"), "dZcat", k.get("gemm_nt_dZcat"))
PY
}
run WN_BWD3=1
run WN_BWD3=0
