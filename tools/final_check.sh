timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
WN_BWD3=0 timeout 200 python -m pytest tests/test_gpu_fast.py -m gpu -x -q 2>&1 | tail -1
WN_BWD_UNFUSED=1 timeout 200 python -m pytest tests/test_gpu_fast.py -m gpu -x -q 2>&1 | tail -1
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["step_tensor_frac"])
print(d["generation"]["samples_per_s_per_stream"], d["autoencoder"]["samples_per_s"], d["cpu_baseline"]["value"])
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
