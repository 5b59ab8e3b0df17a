run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gen-steps 0 --no-ae > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$*" <<'PY'
import json, sys
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
k={x["kernel"]: round(x["ms_per_step"], 3) for x in d["kernels"]}
print(sys.argv[1], d["ms_per_step"], "bwd2", k.get("block_bwd2"), "fwd", k.get("block_fwd"), "dx", k.get("gemm_nt_dx"), "dZcat", k.get("gemm_nt_dZcat"))
PY
}
run WN_BWD3_DBG=0
run WN_BWD3_DBG=1
run WN_BWD3_DBG=2
run WN_BWD3_DBG=3
run WN_BWD3=0
