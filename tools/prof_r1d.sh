# Captures of the build with the three-stage block backward kernel (default): full GPU test suite, smoke, launch list of one
# warm train step, one --set full capture of block_bwd3, then the bench.
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
STEPS=2 timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv python tools/one_step.py > gpurun_out/r1d_one_step.log 2>&1
STEPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:block_bwd3 -s 35 -c 1 -o gpurun_out/r1d_block_bwd3 -f python tools/one_step.py >> gpurun_out/r1d_one_step.log 2>&1
tail -2 gpurun_out/r1d_one_step.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["step_tensor_frac"])
print(d["generation"]["samples_per_s_per_stream"], d["autoencoder"]["samples_per_s"], d["cpu_baseline"]["value"])
PY
