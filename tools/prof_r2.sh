# Round-2 captures: full GPU test suite, smoke, launch list of one warm train step, --set full captures of the fused block backward,
# the block forward, the pipelined generation kernel and one encoder layer of the autoencoder, then the default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest.log
tail -1 gpurun_out/r2_pytest.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
STEPS=2 timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/one_step.py > gpurun_out/r2_one_step.log 2>&1
STEPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:block_bwd6 -s 35 -c 1 -o gpurun_out/r2_block_bwd6 -f python tools/one_step.py >> gpurun_out/r2_one_step.log 2>&1
STEPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:block_fwd2 -s 35 -c 1 -o gpurun_out/r2_block_fwd2 -f python tools/one_step.py >> gpurun_out/r2_one_step.log 2>&1
STEPS=200 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gen_pipe -s 1 -c 1 -o gpurun_out/r2_gen_pipe -f python tools/gen_run.py >> gpurun_out/r2_one_step.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"enc_|cond_|frame_sum|block_" -s 600 -c 400 --csv --log-file gpurun_out/r2_ae_launches.csv python tools/ae_profile.py auto > gpurun_out/r2_ae_step.log 2>&1
tail -2 gpurun_out/r2_one_step.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["step_tensor_frac"])
print(d["generation"]["us_per_step"], d["generation"]["kernel"], d["autoencoder"]["ms_per_step"], d["autoencoder"]["tflops"], d["cpu_baseline"]["value"])
PY
